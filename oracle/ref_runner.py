"""
oracle/ref_runner.py -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

Runs the UNMODIFIED reference (homonim 0.4.3) on in-memory rasters, the way its own `RasterFuse.process` drives it:

  * `load()` imports the reference's `homonim.kernel_model` / `raster_array` / `enums` from `baseline/_ref` (where
    `__graft_entry__.build()` pip-installs /root/reference -- git-ignored, travels to the GPU box) or, failing that,
    from /root/reference itself, through the `oracle/rasterio_stub` name stub.  Everything the kernel-model path does
    in numpy / cv2 is the reference's real code; its two GDAL calls (`rasterio.warp.reproject`,
    `rasterio.fill.fillnodata`) are served by the C restatement in oracle/gdal_restate.* (rasterio / GDAL cannot be
    installed here -- see DESIGN.md section 2).
  * `fuse_band` = one (band, block) of `RasterFuse._process_block` (fuse.py:295-319) with ONE block: the reference's
    `max_block_mem -> inf` result, which is what the GPU path computes.
  * `fuse_band_blocked` = the reference's own default mode: the block grid of `RasterPairReader.block_pairs`
    (raster_pair.py:227-269 `_auto_block_shape`, :379-428) for `max_block_mem` MB with `overlap_for_kernel` overlap
    (utils.py:136-153), the blocks run on a `ThreadPoolExecutor(max_workers=threads)` as in fuse.py:396-408.  The file
    reads / writes of `_process_block` become array slicing; windows in the "other" image follow
    `utils.expand_window_to_grid` / `round_window_to_grid` (utils.py:59-103).

Used by `bench.py --impl reference` and the bench's `cpu_baseline` leg only.
"""
import importlib
import math
import pathlib
import sys
from concurrent.futures import ThreadPoolExecutor
from types import SimpleNamespace

import numpy as np

_REPO = pathlib.Path(__file__).resolve().parent.parent
_STUB_DIR = pathlib.Path(__file__).resolve().parent / 'rasterio_stub'
CANDIDATES = (_REPO / 'baseline' / '_ref', pathlib.Path('/root/reference'))
NAN = float('nan')
_cached = None


def load():
    """ The reference's modules (namespace with kernel_model, raster_array, enums, utils, rasterio, root) or None. """
    global _cached
    if _cached is not None:
        return _cached or None
    root = next((c for c in CANDIDATES if (c / 'homonim' / 'kernel_model.py').exists()), None)
    if root is None:
        _cached = False
        return None
    for p in (str(_REPO), str(_STUB_DIR), str(root)):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        rasterio = importlib.import_module('rasterio')
        if getattr(rasterio, '__version__', '') != '0.0-oracle-stub':
            raise ImportError('a real rasterio is importable; the stub is not needed')
        ns = SimpleNamespace(
            kernel_model=importlib.import_module('homonim.kernel_model'),
            raster_array=importlib.import_module('homonim.raster_array'),
            enums=importlib.import_module('homonim.enums'),
            utils=importlib.import_module('homonim.utils'),
            rasterio=rasterio, root=str(root),
        )
    except Exception:           # pragma: no cover  (missing dependency of the reference: fall back to the port)
        _cached = False
        return None
    _cached = ns
    return ns


def _affine(ns, t):
    return ns.rasterio.Affine(*t[:6])


def _model(ns, model, kernel_shape, proc_crs, r2_inpaint_thresh, find_r2=False):
    km = ns.kernel_model
    cls = km.SrcSpaceModel if proc_crs == 'src' else km.RefSpaceModel          # fuse.py:376
    return cls(model=ns.enums.Model(model), kernel_shape=tuple(kernel_shape), find_r2=find_r2,
               r2_inpaint_thresh=r2_inpaint_thresh)


def _fit_apply(ns, kmodel, src_blk, src_tf, src_nodata, ref_blk, ref_tf):
    """ fuse.py:304-307 on one block: arrays as the reference's reader delivers them (float32, raster_array.py:183-188). """
    RasterArray = ns.raster_array.RasterArray
    crs = ns.rasterio.CRS.from_epsg(32735)
    src32 = np.asarray(src_blk, dtype='float32')
    mk = lambda: RasterArray(src32.copy(), crs, src_tf, nodata=src_nodata)   # noqa: E731  (fit mutates its inputs)
    param_ra = kmodel.fit(mk(), RasterArray(np.array(ref_blk, dtype='float32'), crs, ref_tf, nodata=NAN))
    corr_ra = kmodel.apply(RasterArray(src32, crs, src_tf, nodata=src_nodata), param_ra)
    return param_ra, corr_ra


def _windows_one_block(ns, src, src_tf, ref, ref_tf):
    """ raster_pair.py:292-296: reference window covering the source, source window covering that (boundless). """
    hs, ws = src.shape
    inv = ~ref_tf
    c0, r0 = inv * (src_tf * (0, 0))
    c1, r1 = inv * (src_tf * (ws, hs))
    rc0, rr0 = max(int(math.floor(c0)), 0), max(int(math.floor(r0)), 0)
    rc1, rr1 = min(int(math.ceil(c1)), ref.shape[1]), min(int(math.ceil(r1)), ref.shape[0])
    ref_blk_tf = ref_tf * ns.rasterio.Affine.translation(rc0, rr0)
    sinv = ~src_tf
    sc0, sr0 = sinv * (ref_blk_tf * (0, 0))
    sc1, sr1 = sinv * (ref_blk_tf * (rc1 - rc0, rr1 - rr0))
    pc0, pr0 = min(int(math.floor(sc0 + 1e-9)), 0), min(int(math.floor(sr0 + 1e-9)), 0)
    pc1, pr1 = max(int(math.ceil(sc1 - 1e-9)), ws), max(int(math.ceil(sr1 - 1e-9)), hs)
    return (rr0, rr1, rc0, rc1), (pr0, pr1, pc0, pc1)


def _read_boundless(arr, r0, r1, c0, c1, fill, dtype='float32'):
    """ RasterArray.from_rio_dataset's boundless block read (raster_array.py:175-199) on an array. """
    h, w = arr.shape
    out = np.full((r1 - r0, c1 - c0), fill, dtype=dtype)
    a0, a1, b0, b1 = max(r0, 0), min(r1, h), max(c0, 0), min(c1, w)
    if a1 > a0 and b1 > b0:
        out[a0 - r0:a1 - r0, b0 - c0:b1 - c0] = arr[a0:a1, b0:b1]
    return out


def fuse_band(ns, src, src_transform, src_nodata, ref, ref_transform, model, kernel_shape, proc_crs='ref',
              r2_inpaint_thresh=0.25, find_r2=False):
    """ One band, ONE block: returns (params [2|3, h, w] on the proc grid window, corr [hs, ws]).  ``find_r2`` is what
    `RasterFuse.process` passes when a parameter image is asked for (fuse.py:377). """
    src_tf, ref_tf = _affine(ns, src_transform), _affine(ns, ref_transform)
    (rr0, rr1, rc0, rc1), (pr0, pr1, pc0, pc1) = _windows_one_block(ns, src, src_tf, ref, ref_tf)
    fill = NAN if src_nodata is None else src_nodata
    src_blk = _read_boundless(src, pr0, pr1, pc0, pc1, fill)
    ref_blk = ref[rr0:rr1, rc0:rc1]
    T = ns.rasterio.Affine.translation
    kmodel = _model(ns, model, kernel_shape, proc_crs, r2_inpaint_thresh, find_r2)
    param_ra, corr_ra = _fit_apply(ns, kmodel, src_blk, src_tf * T(pc0, pr0), src_nodata, ref_blk, ref_tf * T(rc0, rr0))
    hs, ws = src.shape
    corr = np.ascontiguousarray(corr_ra.array[-pr0:-pr0 + hs, -pc0:-pc0 + ws])
    params = param_ra.array
    if proc_crs == 'src':
        params = np.ascontiguousarray(params[:, -pr0:-pr0 + hs, -pc0:-pc0 + ws])
    return params, corr


def auto_block_shape(proc_shape, src_res, ref_res, proc_crs, max_block_mem):
    """ raster_pair.py:227-269 `_auto_block_shape`. """
    src_area, ref_area = float(np.prod(np.abs(src_res))), float(np.prod(np.abs(ref_res)))
    if proc_crs == 'ref':
        mem_scale = src_area / ref_area if ref_area > src_area else 1.0
    else:
        mem_scale = 1.0 if ref_area > src_area else ref_area / src_area
    limit = (max_block_mem * mem_scale if max_block_mem > 0 else np.inf) * 2 ** 20
    shape = np.array(proc_shape, dtype='float')
    while np.prod(shape) * 4 > limit:
        shape[np.argmax(shape)] /= 2
    return tuple(int(v) for v in np.ceil(shape))


def fuse_band_blocked(ns, src, src_transform, src_nodata, ref, ref_transform, model, kernel_shape, proc_crs='ref',
                      r2_inpaint_thresh=0.25, max_block_mem=100.0, threads=1, executor=None):
    """
    One band through the reference's block grid.  Returns (corr [hs, ws], number of blocks).  ``executor``: a shared
    ThreadPoolExecutor (the reference runs the blocks of ALL bands on one pool); otherwise one is made for this band.
    """
    src_tf, ref_tf = _affine(ns, src_transform), _affine(ns, ref_transform)
    T = ns.rasterio.Affine.translation
    hs, ws = src.shape
    (rr0, rr1, rc0, rc1), _ = _windows_one_block(ns, src, src_tf, ref, ref_tf)
    kh, kw = kernel_shape
    overlap = (int(math.ceil(kh / 2)), int(math.ceil(kw / 2)))               # utils.overlap_for_kernel
    if proc_crs == 'ref':
        p_r0, p_r1, p_c0, p_c1 = rr0, rr1, rc0, rc1
        proc_tf, other_tf = ref_tf, src_tf
    else:
        p_r0, p_r1, p_c0, p_c1 = 0, hs, 0, ws
        proc_tf, other_tf = src_tf, ref_tf
    block_shape = auto_block_shape((p_r1 - p_r0, p_c1 - p_c0), (src_tf.a, src_tf.e), (ref_tf.a, ref_tf.e), proc_crs,
                                   max_block_mem)
    if block_shape[0] <= overlap[0] or block_shape[1] <= overlap[1]:
        raise ValueError('The auto block shape is smaller than the overlap.  Increase `max_block_mem`.')
    fill = NAN if src_nodata is None else src_nodata
    corr_out = np.full((hs, ws), NAN, dtype='float32')
    kmodel = _model(ns, model, kernel_shape, proc_crs, r2_inpaint_thresh)
    oinv = ~other_tf

    def other_window(r0, r1, c0, c1, expand):
        """ window of the other image covering proc window rows [r0, r1) x cols [c0, c1) """
        x0, y0 = proc_tf * (c0, r0)
        x1, y1 = proc_tf * (c1, r1)
        oc0, or0 = oinv * (x0, y0)
        oc1, or1 = oinv * (x1, y1)
        if expand:
            return int(math.floor(or0 + 1e-9)), int(math.ceil(or1 - 1e-9)), int(math.floor(oc0 + 1e-9)), \
                int(math.ceil(oc1 - 1e-9))
        return int(round(or0)), int(round(or1)), int(round(oc0)), int(round(oc1))

    jobs = []
    for ul_r in range(p_r0 - overlap[0], p_r1 - overlap[0], block_shape[0]):
        for ul_c in range(p_c0 - overlap[1], p_c1 - overlap[1], block_shape[1]):
            br_r, br_c = ul_r + block_shape[0] + 2 * overlap[0], ul_c + block_shape[1] + 2 * overlap[1]
            in_w = (max(ul_r, p_r0), min(br_r, p_r1), max(ul_c, p_c0), min(br_c, p_c1))
            out_w = (max(ul_r + overlap[0], p_r0), min(br_r - overlap[0], p_r1), max(ul_c + overlap[1], p_c0),
                     min(br_c - overlap[1], p_c1))
            jobs.append((in_w, out_w))

    def run(job):
        (ir0, ir1, ic0, ic1), (or0, or1, oc0, oc1) = job
        if proc_crs == 'ref':
            s_r0, s_r1, s_c0, s_c1 = other_window(ir0, ir1, ic0, ic1, True)          # source in-block (boundless)
            src_blk = _read_boundless(src, s_r0, s_r1, s_c0, s_c1, fill)
            ref_blk = ref[ir0:ir1, ic0:ic1]
            _, corr_ra = _fit_apply(ns, kmodel, src_blk, src_tf * T(s_c0, s_r0), src_nodata, ref_blk,
                                    ref_tf * T(ic0, ir0))
            w_r0, w_r1, w_c0, w_c1 = other_window(or0, or1, oc0, oc1, False)        # source out-block
        else:
            src_blk = src[ir0:ir1, ic0:ic1]
            f_r0, f_r1, f_c0, f_c1 = other_window(ir0, ir1, ic0, ic1, True)          # reference in-block
            f_r0, f_c0 = max(f_r0, 0), max(f_c0, 0)
            f_r1, f_c1 = min(f_r1, ref.shape[0]), min(f_c1, ref.shape[1])
            ref_blk = ref[f_r0:f_r1, f_c0:f_c1]
            _, corr_ra = _fit_apply(ns, kmodel, src_blk, src_tf * T(ic0, ir0), src_nodata, ref_blk,
                                    ref_tf * T(f_c0, f_r0))
            s_r0, s_c0 = ir0, ic0
            w_r0, w_r1, w_c0, w_c1 = or0, or1, oc0, oc1
        w_r0, w_c0, w_r1, w_c1 = max(w_r0, 0), max(w_c0, 0), min(w_r1, hs), min(w_c1, ws)
        if w_r1 > w_r0 and w_c1 > w_c0:                                              # fuse.py:309-313 (write)
            corr_out[w_r0:w_r1, w_c0:w_c1] = corr_ra.array[w_r0 - s_r0:w_r1 - s_r0, w_c0 - s_c0:w_c1 - s_c0]

    if executor is not None:
        return corr_out, [executor.submit(run, j) for j in jobs]
    if threads <= 1:
        for j in jobs:
            run(j)
    else:
        with ThreadPoolExecutor(max_workers=threads) as pool:
            list(pool.map(run, jobs))
    return corr_out, len(jobs)

"""
oracle/make_golden.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Regenerates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/homonim, imported through
oracle/rasterio_stub) on small seeded inputs.  Runs only in the build container (the GPU box has no
/root/reference); the .npz fixtures it writes are committed and are what pins oracle/kernel_model_np.py (CPU tests)
and what the CUDA path is compared with at run time (GPU tests).

    python -m oracle.make_golden

Every fixture stores the inputs (arrays, 6-coefficient transforms, nodata), the configuration, and the reference's
outputs: ``params`` (fit) and ``corr`` (apply).  GDAL-backed steps inside the reference (reproject / fillnodata) are
served by oracle/gdal_restate.py -- see its header for the parity caveat.
"""
import json
import pathlib
import sys
import warnings

import numpy as np

REPO = pathlib.Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

from oracle.ref_import import import_reference  # noqa: E402

GOLDEN_DIR = REPO / 'tests' / 'golden'
NAN = float('nan')


def _texture(rng, h, w, sigma=2.0):
    import cv2
    t = cv2.GaussianBlur(rng.standard_normal((h, w)).astype('float32'), (0, 0), sigma)
    return (t / t.std()).astype('float32')


def same_grid_inputs(seed, h=48, w=40, src_nodata=NAN, ref_nodata=NAN, mu=3000.0):
    rng = np.random.default_rng(seed)
    t = _texture(rng, h, w)
    src = (mu + 0.3 * mu * t + 0.02 * mu * rng.standard_normal((h, w))).astype('float32')
    yy, xx = np.mgrid[0:h, 0:w].astype('float32')
    gain = 0.6 + 0.2 * np.sin(xx / 9.0 + yy / 17.0)
    off = 0.05 * mu * np.cos(yy / 11.0)
    ref = (gain * src + off + 0.01 * mu * rng.standard_normal((h, w))).astype('float32')
    ref[20:27, 14:22] = (mu * (0.5 + 0.3 * rng.standard_normal((7, 8)))).astype('float32')   # low R2 blob
    src_bad = np.zeros((h, w), bool)
    src_bad[:6, :5] = True
    src_bad[30:34, 8:13] = True
    src_bad[rng.integers(0, h, 12), rng.integers(0, w, 12)] = True
    ref_bad = np.zeros((h, w), bool)
    ref_bad[-4:, -7:] = True
    ref_bad[10:12, 30:33] = True
    src[src_bad] = src_nodata
    ref[ref_bad] = ref_nodata
    return src, ref


def upsample_pattern(rng, hp, wp, ratio, mu, dtype, src_nodata):
    """ hi-res source over an hp x wp proc grid, plus its matching coarse reference (fully valid). """
    t = _texture(rng, hp, wp, 2.0)
    lo = mu + 0.3 * mu * t
    import cv2
    hi = cv2.resize(lo, (wp * ratio, hp * ratio), interpolation=cv2.INTER_LINEAR)
    hi = hi + 0.02 * mu * rng.standard_normal(hi.shape).astype('float32')
    if dtype != 'float32':
        hi = np.clip(np.round(hi), 1, np.iinfo(dtype).max)
    avg = hi.reshape(hp, ratio, wp, ratio).mean(axis=(1, 3))
    yy, xx = np.mgrid[0:hp, 0:wp].astype('float32')
    ref = ((0.6 + 0.2 * np.sin(xx / 7.0 + yy / 13.0)) * avg + 0.05 * mu * np.cos(yy / 9.0)
           + 0.01 * mu * rng.standard_normal((hp, wp))).astype('float32')
    ref[12:18, 9:16] = (mu * (0.5 + 0.3 * rng.standard_normal((6, 7)))).astype('float32')
    bad = np.zeros((hp, wp), bool)
    bad[:5, :4] = True
    bad[22:26, 18:23] = True
    bad_hi = np.kron(bad, np.ones((ratio, ratio), bool)).astype(bool)
    hi = hi.astype(dtype)
    hi[bad_hi] = src_nodata
    return hi, ref


def main():
    km, ra_mod, enums, rio = import_reference()
    Affine, CRS = rio.Affine, rio.crs.CRS
    RasterArray = ra_mod.RasterArray
    crs = CRS({'init': 'epsg:3857'})
    GOLDEN_DIR.mkdir(parents=True, exist_ok=True)
    warnings.simplefilter('ignore')
    index = {}

    def save(name, meta, **arrays):
        np.savez_compressed(GOLDEN_DIR / f'{name}.npz', **arrays)
        index[name] = meta
        print(f'{name}: ' + ', '.join(f'{k}{tuple(v.shape)}' for k, v in arrays.items()))

    # ---- A. same-grid KernelModel.fit / apply -------------------------------------------------------------------
    tf = Affine(10, 0, 1000, 0, -10, 5000)
    same_grid_cases = [
        ('gain', (1, 1), False, None, NAN), ('gain', (3, 3), True, None, NAN), ('gain', (5, 7), False, None, 0.0),
        ('gain-blk-offset', (1, 1), False, None, NAN), ('gain-blk-offset', (5, 5), True, None, NAN),
        ('gain-blk-offset', (15, 15), False, None, 0.0),
        ('gain-offset', (5, 5), False, None, NAN), ('gain-offset', (5, 7), True, None, NAN),
        ('gain-offset', (15, 15), True, None, 0.0), ('gain-offset', (5, 5), False, 0.25, NAN),
        ('gain-offset', (7, 7), True, 0.5, NAN),
    ]
    for ci, (model, kshape, find_r2, thresh, src_nodata) in enumerate(same_grid_cases):
        src, ref = same_grid_inputs(100 + ci, src_nodata=src_nodata)
        kmodel = km.KernelModel(model, kshape, find_r2=find_r2, r2_inpaint_thresh=thresh)
        src_ra = RasterArray(src.copy(), crs, tf, nodata=src_nodata)
        ref_ra = RasterArray(ref.copy(), crs, tf, nodata=NAN)
        params = kmodel.fit(src_ra, ref_ra).array
        corr = kmodel.apply(RasterArray(src.copy(), crs, tf, nodata=src_nodata),
                            RasterArray(params.copy(), crs, tf, nodata=NAN)).array
        name = f'same_{ci:02d}_{model}_k{kshape[0]}x{kshape[1]}' + ('_r2' if find_r2 else '') + \
               (f'_inp{thresh}' if thresh is not None else '')
        save(name, dict(kind='same', model=model, kernel_shape=kshape, find_r2=find_r2, r2_inpaint_thresh=thresh,
                        src_nodata=None if src_nodata is None else float(src_nodata), ref_nodata=NAN,
                        transform=list(tf)),
             src=src, ref=ref, params=params, corr=corr)

    # ---- B. RefSpaceModel: fit on the reference grid, apply on the source grid ------------------------------------
    ref_tf = Affine(10, 0, 2000, 0, -10, 9000)
    ref_cases = [
        # model, kernel, find_r2, thresh, ratio, dtype, src_nodata, (col, row) source offset in source pixels, partial
        ('gain', (1, 1), False, None, 4, 'uint16', 0, (0, 0), False),
        ('gain-blk-offset', (5, 5), True, None, 4, 'uint16', 0, (0, 0), False),
        ('gain-offset', (5, 5), False, 0.25, 4, 'uint16', 0, (0, 0), False),
        ('gain-offset', (7, 7), True, None, 5, 'float32', NAN, (0, 0), False),
        ('gain-blk-offset', (3, 5), False, None, 4, 'uint8', 0, (0, 0), True),
        ('gain-blk-offset', (5, 5), False, None, 4, 'float32', NAN, (1.3, 2.6), False),   # mis-aligned grids
        ('gain-offset', (5, 5), True, 0.25, 3, 'uint16', 0, (0.5, 0.25), True),
    ]
    for ci, (model, kshape, find_r2, thresh, ratio, dtype, src_nodata, shift, partial) in enumerate(ref_cases):
        rng = np.random.default_rng(200 + ci)
        hp, wp = 36, 30
        mu = 120.0 if dtype == 'uint8' else 3000.0
        src, ref = upsample_pattern(rng, hp, wp, ratio, mu, dtype, src_nodata)
        # pad the reference by 2 pixels (fully valid, replicated) so that it encompasses the source
        ref = np.pad(ref, 2, mode='edge')
        ref_full_tf = ref_tf * Affine.translation(-2, -2)
        src_tf = ref_tf * Affine.scale(1.0 / ratio) * Affine.translation(*shift)
        # reference window covering the source, source window covering that (raster_pair.py:292-296), boundless read
        # of the source (raster_array.py:175-199): done here by explicit cropping / nodata padding
        hs, ws = src.shape
        inv = ~ref_full_tf
        c0, r0 = inv * (src_tf * (0, 0))
        c1, r1 = inv * (src_tf * (ws, hs))
        rc0, rr0, rc1, rr1 = int(np.floor(c0)), int(np.floor(r0)), int(np.ceil(c1)), int(np.ceil(r1))
        ref_blk = ref[rr0:rr1, rc0:rc1].copy()
        ref_blk_tf = ref_full_tf * Affine.translation(rc0, rr0)
        sinv = ~src_tf
        sc0, sr0 = sinv * (ref_blk_tf * (0, 0))
        sc1, sr1 = sinv * (ref_blk_tf * (ref_blk.shape[1], ref_blk.shape[0]))
        pc0, pr0 = int(np.floor(sc0 + 1e-9)), int(np.floor(sr0 + 1e-9))
        pc1, pr1 = int(np.ceil(sc1 - 1e-9)), int(np.ceil(sr1 - 1e-9))
        fill = src_nodata
        src_blk = np.full((pr1 - pr0, pc1 - pc0), fill, dtype='float32')
        src_blk[-pr0:-pr0 + hs, -pc0:-pc0 + ws] = src.astype('float32')
        src_blk_tf = src_tf * Affine.translation(pc0, pr0)

        kmodel = km.RefSpaceModel(model, kshape, find_r2=find_r2, r2_inpaint_thresh=thresh, mask_partial=partial)
        mk = lambda: RasterArray(src_blk.copy(), crs, src_blk_tf, nodata=src_nodata)   # noqa: E731
        param_ra = kmodel.fit(mk(), RasterArray(ref_blk.copy(), crs, ref_blk_tf, nodata=NAN))
        corr_blk = kmodel.apply(mk(), param_ra).array
        corr = corr_blk[-pr0:-pr0 + hs, -pc0:-pc0 + ws]                 # written extent = the source (fuse.py:311)
        name = f'refspace_{ci:02d}_{model}_k{kshape[0]}x{kshape[1]}_r{ratio}_{dtype}' + ('_partial' if partial else '')
        save(name, dict(kind='refspace', model=model, kernel_shape=kshape, find_r2=find_r2, r2_inpaint_thresh=thresh,
                        mask_partial=partial, src_nodata=None if src_nodata is None else float(src_nodata),
                        ref_nodata=NAN, src_transform=list(src_tf), ref_transform=list(ref_full_tf),
                        param_transform=list(ref_blk_tf)),
             src=src, ref=ref, params=param_ra.array, corr=np.ascontiguousarray(corr))

    # ---- C. SrcSpaceModel: fit and apply on the source grid (reference coarser, cubic-spline up-sampled) ----------
    src_cases = [
        ('gain', (3, 3), True, None, False), ('gain-blk-offset', (5, 5), False, None, False),
        ('gain-offset', (5, 5), True, None, False), ('gain-offset', (7, 5), False, 0.25, False),
        ('gain-blk-offset', (3, 3), False, None, True),
    ]
    for ci, (model, kshape, find_r2, thresh, partial) in enumerate(src_cases):
        rng = np.random.default_rng(300 + ci)
        hp, wp, ratio = 30, 26, 2
        src, ref = upsample_pattern(rng, hp, wp, ratio, 0.3, 'float32', NAN)
        ref[3:5, 20:23] = NAN                                           # a hole in the reference
        src_tf = ref_tf * Affine.scale(1.0 / ratio)
        kmodel = km.SrcSpaceModel(model, kshape, find_r2=find_r2, r2_inpaint_thresh=thresh, mask_partial=partial)
        param_ra = kmodel.fit(RasterArray(src.copy(), crs, src_tf, nodata=NAN),
                              RasterArray(ref.copy(), crs, ref_tf, nodata=NAN))
        corr = kmodel.apply(RasterArray(src.copy(), crs, src_tf, nodata=NAN), param_ra).array
        name = f'srcspace_{ci:02d}_{model}_k{kshape[0]}x{kshape[1]}' + ('_partial' if partial else '')
        save(name, dict(kind='srcspace', model=model, kernel_shape=kshape, find_r2=find_r2, r2_inpaint_thresh=thresh,
                        mask_partial=partial, src_nodata=NAN, ref_nodata=NAN, src_transform=list(src_tf),
                        ref_transform=list(ref_tf)),
             src=src, ref=ref, params=param_ra.array, corr=corr)

    # ---- D. the reference's own conftest fixtures (tests/conftest.py:74-89, 112-140) ------------------------------
    a100 = np.array(range(1, 201), dtype='float32').reshape(20, 10)
    a100[:, [0, -1]] = NAN
    a100[[0, -1], :] = NAN
    a50 = np.kron(a100, np.ones((2, 2))).astype('float32')
    a50[:, [0, 1, -2, -1]] = NAN
    a50[[0, 1, -2, -1], :] = NAN
    tf100 = Affine(1, 0, 0, 0, -1, 0) * Affine.translation(5, 5)
    tf50 = tf100 * Affine.scale(0.5)
    for model, kshape in [('gain', (3, 3)), ('gain-blk-offset', (5, 5)), ('gain-offset', (5, 5))]:
        kmodel = km.RefSpaceModel(model, kshape, mask_partial=False, r2_inpaint_thresh=0.25)
        param_ra = kmodel.fit(RasterArray(a50.copy(), crs, tf50, nodata=NAN),
                              RasterArray(a100.copy(), crs, tf100, nodata=NAN))
        corr = kmodel.apply(RasterArray(a50.copy(), crs, tf50, nodata=NAN), param_ra).array
        save(f'conftest_ref_{model}_k{kshape[0]}x{kshape[1]}',
             dict(kind='refspace', model=model, kernel_shape=kshape, find_r2=False, r2_inpaint_thresh=0.25,
                  mask_partial=False, src_nodata=NAN, ref_nodata=NAN, src_transform=list(tf50),
                  ref_transform=list(tf100), param_transform=list(tf100)),
             src=a50, ref=a100, params=param_ra.array, corr=corr)

    def _clean(meta):
        return {k: (None if isinstance(v, float) and np.isnan(v) else v) for k, v in meta.items()}

    # NaN nodata is stored as the string 'nan' (JSON has no NaN)
    for name, meta in index.items():
        for key in ('src_nodata', 'ref_nodata'):
            if isinstance(meta.get(key), float) and np.isnan(meta[key]):
                meta[key] = 'nan'
        if isinstance(meta.get('r2_inpaint_thresh'), float) and np.isinf(meta['r2_inpaint_thresh']):
            meta['r2_inpaint_thresh'] = '-inf'
    (GOLDEN_DIR / 'index.json').write_text(json.dumps(index, indent=1, sort_keys=True))
    print(f'{len(index)} fixtures written to {GOLDEN_DIR}')


if __name__ == '__main__':
    main()

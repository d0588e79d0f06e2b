"""
GPU tier: parity of the sm_100a CUDA path against (i) the golden vectors generated from the unmodified reference and
(ii) the oracle port on larger seeded synthetic rasters.  Tolerance (BASELINE.json north_star): parameter and
corrected-pixel max rel err <= 1e-4 (see conftest.check_params for the floors), nodata masks identical.
"""
import warnings

import numpy as np
import pytest

from conftest import RTOL, assert_same_mask, check_corr, check_params, golden_names, load_golden, rel_err

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from homonim_b200 import (CRS, KernelModel, Model, ProcCrs, RasterArray, RasterFuse, RefSpaceModel,   # noqa: E402
                          SrcSpaceModel)
from homonim_b200.synthetic import make_pair   # noqa: E402

CRS0 = CRS.from_epsg(3857)


def _model_kwargs(meta):
    kw = dict(find_r2=meta['find_r2'], r2_inpaint_thresh=meta['r2_inpaint_thresh'])
    if 'mask_partial' in meta:
        kw['mask_partial'] = meta['mask_partial']
    return kw


@pytest.mark.parametrize('device_resident', [False, True])
@pytest.mark.parametrize('name', golden_names('same'))
def test_same_grid_vs_reference_golden(name, device_resident):
    meta, g = load_golden(name)
    model = KernelModel(meta['model'], meta['kernel_shape'], **_model_kwargs(meta))
    src_ra = RasterArray(g['src'].copy(), CRS0, meta['transform'], nodata=meta['src_nodata'])
    ref_ra = RasterArray(g['ref'].copy(), CRS0, meta['transform'], nodata=meta['ref_nodata'])
    if device_resident:
        src_ra, ref_ra = src_ra.to_device(), ref_ra.to_device()
    param_ra = model.fit(src_ra, ref_ra)
    assert param_ra.is_device == device_resident
    # inputs are not mutated
    assert np.array_equal(src_ra.to_host().array, g['src'], equal_nan=True)
    corr_ra = model.apply(src_ra, param_ra)
    params, corr = param_ra.to_host().array, corr_ra.to_host().array
    valid = g['src'][~np.isnan(g['params'][0])]
    check_params(params, g['params'], float(np.mean(valid)) if valid.size else 1.0, name)
    check_corr(corr, g['corr'], name)
    assert param_ra.transform == src_ra.transform and np.isnan(param_ra.nodata)


@pytest.mark.parametrize('name', golden_names('refspace'))
def test_refspace_vs_reference_golden(name):
    meta, g = load_golden(name)
    src_ra = RasterArray(g['src'].copy(), CRS0, meta['src_transform'], nodata=meta['src_nodata'])
    ref_ra = RasterArray(g['ref'].copy(), CRS0, meta['ref_transform'], nodata=meta['ref_nodata'])
    fuse = RasterFuse(src_ra, ref_ra, proc_crs=ProcCrs.ref)
    model = RefSpaceModel(meta['model'], meta['kernel_shape'], **_model_kwargs(meta))
    corr_ra, param_ra = fuse._process_band(0, model)
    assert np.allclose(tuple(param_ra.transform), meta['param_transform'])
    src_f = g['src'].astype('float32')
    check_params(param_ra.array, g['params'], float(np.mean(src_f[np.isfinite(src_f) & (src_f != 0)])), name)
    check_corr(corr_ra.array, g['corr'].astype('float32'), name)
    assert corr_ra.shape == src_ra.shape and corr_ra.transform == src_ra.transform


@pytest.mark.parametrize('name', golden_names('srcspace'))
def test_srcspace_vs_reference_golden(name):
    meta, g = load_golden(name)
    src_ra = RasterArray(g['src'].copy(), CRS0, meta['src_transform'], nodata=meta['src_nodata'])
    ref_ra = RasterArray(g['ref'].copy(), CRS0, meta['ref_transform'], nodata=meta['ref_nodata'])
    model = SrcSpaceModel(meta['model'], meta['kernel_shape'], **_model_kwargs(meta))
    param_ra = model.fit(src_ra, ref_ra)
    corr_ra = model.apply(src_ra, param_ra)
    check_params(param_ra.array, g['params'], float(np.nanmean(g['src'])), name)
    check_corr(corr_ra.array, g['corr'], name)


# ---- larger seeded rasters against the oracle port ------------------------------------------------------------------
def _oracle():
    from oracle import kernel_model_np as kmnp
    return kmnp


@pytest.mark.parametrize('model, kernel_shape, find_r2, thresh', [
    (Model.gain, (1, 1), False, None),
    (Model.gain, (7, 7), True, None),
    (Model.gain_blk_offset, (5, 5), True, None),
    (Model.gain_blk_offset, (15, 15), False, None),
    (Model.gain_offset, (15, 15), True, None),
    (Model.gain_offset, (31, 31), False, None),
    (Model.gain_offset, (15, 15), False, 0.25),
])
def test_same_grid_1k_vs_oracle(model, kernel_shape, find_r2, thresh):
    """ 1000 x 1200 float32 planes on one grid (the SrcSpace / C3 regime), incl. a width that is not a multiple of 4. """
    kmnp = _oracle()
    src_ra, ref_ra = make_pair(1000, 1203, 1, bands=1, dtype='float32', mu=0.3, seed=11, device='cuda',
                               src_nodata=float('nan'), ref_pad=0)
    src_ra = RasterArray(src_ra.array[0].contiguous(), src_ra.crs, src_ra.transform, nodata=src_ra.nodata)
    ref_ra = RasterArray(ref_ra.array[0].contiguous(), ref_ra.crs, src_ra.transform, nodata=ref_ra.nodata)
    km = KernelModel(model, kernel_shape, find_r2=find_r2, r2_inpaint_thresh=thresh)
    param_ra = km.fit(src_ra, ref_ra)
    corr_ra = km.apply(src_ra, param_ra)
    src, ref = src_ra.to_host().array, ref_ra.to_host().array
    exp_params = kmnp.fit_same_grid(src, float('nan'), ref, float('nan'), model, kernel_shape, find_r2, thresh)
    exp_corr = kmnp.apply_same_grid(src, exp_params)
    bad = check_params(param_ra.to_host().array, exp_params, float(np.nanmean(src)), 'params')
    check_corr(corr_ra.to_host().array, exp_corr, 'corr', bad=bad)
    # most pixels are bit-identical, not merely within tolerance (gain-blk-offset too: the block statistics replay numpy's
    # float32 pairwise np.std to the bit)
    got = param_ra.to_host().array[0]
    same = (got == exp_params[0]) | (np.isnan(got) & np.isnan(exp_params[0]))
    assert same.mean() > 0.99


@pytest.mark.parametrize('shape', [(300, 700), (2100, 2300)])
@pytest.mark.parametrize('model, kernel_shape, find_r2', [
    (Model.gain, (1, 67), False),
    (Model.gain_offset, (3, 65), True),
    (Model.gain_offset, (1, 127), False),
    (Model.gain_blk_offset, (67, 1), True),
    (Model.gain, (127, 127), True),
])
def test_same_grid_wide_kernels_vs_oracle(shape, model, kernel_shape, find_r2):
    """
    Kernels wider than a 32-column warp: small rasters (one column per thread) are routed to the 4-column variant above
    31 columns, whose 128-column warps carry kernels up to 127 wide (a window never spans more than two warps).  Small
    and large rasters take different kernel instantiations; both must agree with the oracle.
    """
    kmnp = _oracle()
    h, w = shape
    src_ra, ref_ra = make_pair(h, w, 1, bands=1, dtype='float32', mu=0.3, seed=13, device='cuda',
                               src_nodata=float('nan'), ref_pad=0)
    src_ra = RasterArray(src_ra.array[0].contiguous(), src_ra.crs, src_ra.transform, nodata=src_ra.nodata)
    ref_ra = RasterArray(ref_ra.array[0].contiguous(), ref_ra.crs, src_ra.transform, nodata=ref_ra.nodata)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        km = KernelModel(model, kernel_shape, find_r2=find_r2, r2_inpaint_thresh=None)
    param_ra = km.fit(src_ra, ref_ra)
    src, ref = src_ra.to_host().array, ref_ra.to_host().array
    exp_params = kmnp.fit_same_grid(src, float('nan'), ref, float('nan'), model, kernel_shape, find_r2, None)
    check_params(param_ra.to_host().array, exp_params, float(np.nanmean(src)), f'params {kernel_shape}')


@pytest.mark.parametrize('model, kernel_shape, find_r2', [
    (Model.gain_offset, (31, 31), False), (Model.gain_offset, (15, 5), True), (Model.gain_blk_offset, (15, 15), False),
    (Model.gain_blk_offset, (3, 9), True), (Model.gain, (5, 5), False), (Model.gain, (3, 3), True),
    (Model.gain_offset, (3, 127), False), (Model.gain_offset, (127, 3), False),
])
def test_same_grid_lean_equals_general(model, kernel_shape, find_r2, monkeypatch):
    """ The NaN-nodata form of the fit kernel (csrc/moments.cu, LEAN) performs the same additions in the same order as
    the general form: parameters and the fused corrected image must be bit-identical, on a raster with NaN holes that
    touch every border and whose width is not a multiple of the CTA strip. """
    g = torch.Generator(device='cuda').manual_seed(11)
    h, w = 2100, 2052
    src = torch.rand((h, w), generator=g, device='cuda') * 0.5 + 0.2
    ref = 0.7 * src + 0.05 + 0.02 * torch.rand((h, w), generator=g, device='cuda')
    nan = float('nan')
    for (r0, r1, c0, c1) in ((0, 40, 0, 300), (h - 9, h, 500, 900), (700, 760, w - 33, w), (1000, 1200, 1000, 1003),
                             (300, 301, 0, w), (1500, 1600, 0, 7)):
        src[r0:r1, c0:c1] = nan
    ref[40:80, 1200:1300] = nan
    km = KernelModel(model, kernel_shape, find_r2=find_r2, r2_inpaint_thresh=None)
    norm = km._block_norm(src, nan, ref, nan) if model == Model.gain_blk_offset else None

    def run():
        params = km._fit_planes(src, nan, ref, nan, norm=norm)
        rows = km._fit_planes(src, nan, ref, nan, norm=norm, rows=(311, 1217))
        corr = torch.empty_like(src)
        km._fit_apply_rows(src, nan, ref, nan, 0, h, norm=norm, out=corr)
        torch.cuda.synchronize()
        return params, rows, corr
    lean = run()
    monkeypatch.setenv('HOMONIM_B200_FIT_GENERAL', '1')
    general = run()
    for a, b, what in zip(lean, general, ('params', 'row range', 'fused corr')):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), f'{what}: lean != general'


def test_same_grid_row_range_equals_full_fit():
    """ hb_fit_same_grid_rows / hb_fit_apply_same_grid_rows: the rows of a band inside a larger plane get exactly the
    full fit's values for those rows wherever the band sits (what the row-band shards rely on). """
    src_ra, ref_ra = make_pair(700, 2052, 1, bands=1, dtype='float32', mu=0.3, seed=17, device='cuda',
                               src_nodata=float('nan'), ref_pad=0)
    s, r = src_ra.array[0].contiguous(), ref_ra.array[0].contiguous()
    nan = float('nan')
    for model, kshape, find_r2 in ((Model.gain_offset, (15, 15), True), (Model.gain_blk_offset, (5, 7), False),
                                   (Model.gain, (1, 1), False)):
        km = KernelModel(model, kshape, find_r2=find_r2, r2_inpaint_thresh=None)
        norm = km._block_norm(s, nan, r, nan) if model == Model.gain_blk_offset else None
        full = km._fit_planes(s, nan, r, nan, norm=norm)
        full_corr = km._apply_planes(s, nan, full, mask_src=False)
        for row0, nrows in ((0, 700), (0, 33), (123, 301), (650, 50), (699, 1)):
            part = km._fit_planes(s, nan, r, nan, norm=norm, rows=(row0, nrows))
            exp = full[:, row0:row0 + nrows]
            assert torch.equal(torch.isnan(part), torch.isnan(exp))
            same = ((part == exp) | (torch.isnan(part) & torch.isnan(exp))).float().mean().item()
            assert same > 0.999, (model, row0, nrows, same)
            corr = km._fit_apply_rows(s, nan, r, nan, row0, nrows, norm=norm)
            expc = full_corr[row0:row0 + nrows]
            assert torch.equal(torch.isnan(corr), torch.isnan(expc))
            fin = torch.isfinite(expc)
            rel = ((corr - expc).abs()[fin] / expc.abs()[fin].clamp_min(1e-3 * expc[fin].abs().mean())).max().item()
            assert rel <= 1e-4, (model, row0, nrows, rel)


@pytest.mark.parametrize('dtype, nodata, ratio, model, kernel_shape, thresh', [
    ('uint16', 0, 20, Model.gain, (1, 1), None),
    ('uint16', 0, 20, Model.gain_offset, (15, 15), 0.25),
    ('uint8', 0, 8, Model.gain_blk_offset, (5, 5), None),
    ('float32', float('nan'), 20, Model.gain_blk_offset, (15, 15), None),
    ('float32', float('nan'), 3, Model.gain_offset, (5, 5), None),
])
def test_refspace_fuse_vs_oracle(dtype, nodata, ratio, model, kernel_shape, thresh):
    """ RasterFuse.process (proc_crs=ref) on a 2-band source against the oracle's single-block fuse of each band. """
    kmnp = _oracle()
    hp, wp = 120, 101
    mu = 120.0 if dtype == 'uint8' else (0.3 if dtype == 'float32' else 3000.0)
    src_ra, ref_ra = make_pair(hp, wp, ratio, bands=2, dtype=dtype, mu=mu, seed=5, device='cuda', src_nodata=nodata)
    with RasterFuse(src_ra, ref_ra) as fuse:
        assert fuse.proc_crs == ProcCrs.ref
        corr_ra, param_ra = fuse.process(model=model, kernel_shape=kernel_shape, param_filename='params',
                                         model_config=dict(r2_inpaint_thresh=thresh))
    src, ref = src_ra.to_host(), ref_ra.to_host()
    n_params = param_ra.count // 2
    for b in range(2):
        exp_params, _, exp_corr = kmnp.fuse_band_blocks(
            src.array[b], tuple(src.transform), nodata, ref.array[b], tuple(ref.transform), float('nan'), model,
            kernel_shape, 'ref', True, thresh)
        got_params = np.stack([param_ra.to_host().array[p * 2 + b] for p in range(n_params)])
        valid_src = src.array[b][src.array[b] != nodata] if not np.isnan(nodata) else src.array[b]
        if kernel_shape == (1, 1):
            # R2 of a one-pixel window is 1 - 0/0: both sides return the rounding residue of N*sum(r^2) - sum(r)^2
            # (kernel_model.py:179); require bit-identity almost everywhere instead of a tolerance
            r2_same = (got_params[2] == exp_params[2]) | (np.isnan(got_params[2]) & np.isnan(exp_params[2]))
            assert r2_same.mean() > 0.99
            got_params, exp_params = got_params[:2], exp_params[:2]
        check_params(got_params, exp_params, float(np.nanmean(valid_src.astype('float64'))), f'band {b} params')
        check_corr(corr_ra.to_host().array[b], exp_corr.astype('float32'), f'band {b} corr')


def test_srcspace_fuse_vs_oracle():
    """ proc_crs=src: reference 2x coarser, cubic-spline up-sampled, gain-offset 31x31 (the C3 configuration). """
    kmnp = _oracle()
    src_ra, ref_ra = make_pair(300, 260, 2, bands=1, dtype='float32', mu=0.3, seed=3, device='cuda',
                               src_nodata=float('nan'))
    with RasterFuse(src_ra, ref_ra, proc_crs=ProcCrs.src) as fuse:
        with pytest.warns(Warning):
            RasterFuse(src_ra, ref_ra, proc_crs=ProcCrs.src)
        corr_ra, param_ra = fuse.process(model=Model.gain_offset, kernel_shape=(31, 31), param_filename='p',
                                         model_config=dict(r2_inpaint_thresh=None))
    src, ref = src_ra.to_host(), ref_ra.to_host()
    exp_params, _, exp_corr = kmnp.fuse_band_blocks(
        src.array[0], tuple(src.transform), float('nan'), ref.array[0], tuple(ref.transform), float('nan'),
        Model.gain_offset, (31, 31), 'src', True, None)
    check_params(param_ra.to_host().array, exp_params, float(np.nanmean(src.array[0])), 'params')
    check_corr(corr_ra.to_host().array[0], exp_corr, 'corr')


# ---- reference known-answer tests, run against the CUDA path (tests/test_kernel_model.py of the reference) ----------
def _conftest_rasters():
    a100 = np.array(range(1, 201), dtype='float32').reshape(20, 10)
    a100[:, [0, -1]] = np.nan
    a100[[0, -1], :] = np.nan
    a50 = np.kron(a100, np.ones((2, 2))).astype('float32')
    a50[:, [0, 1, -2, -1]] = np.nan
    a50[[0, 1, -2, -1], :] = np.nan
    from homonim_b200 import Affine
    tf100 = Affine(1, 0, 0, 0, -1, 0) * Affine.translation(5, 5)
    tf50 = tf100 * Affine.scale(0.5)
    return (RasterArray(a100, CRS0, tf100, nodata=float('nan')), RasterArray(a50, CRS0, tf50, nodata=float('nan')))


@pytest.mark.parametrize('model, kernel_shape', [
    (Model.gain, (1, 1)), (Model.gain, (3, 3)), (Model.gain_blk_offset, (1, 1)), (Model.gain_blk_offset, (5, 5)),
    (Model.gain_offset, (5, 5)),
])
def test_ref_and_src_basic_fit(model, kernel_shape):
    """ reference tests/test_kernel_model.py:41-81: same surface at two resolutions => gain ~ 1, offset ~ 0. """
    ra100, ra50 = _conftest_rasters()
    for cls, src_ra, ref_ra in ((RefSpaceModel, ra50, ra100), (SrcSpaceModel, ra100, ra50)):
        km = cls(model, kernel_shape, mask_partial=False, r2_inpaint_thresh=0.25)
        param_ra = km.fit(src_ra, ref_ra.copy())
        proc_ra = ref_ra if cls is RefSpaceModel else src_ra
        assert param_ra.shape == proc_ra.shape and param_ra.transform == proc_ra.transform
        assert (proc_ra.mask == param_ra.mask).all()
        assert param_ra.array[0, param_ra.mask] == pytest.approx(1, abs=1e-2)
        assert param_ra.array[1, param_ra.mask] == pytest.approx(0, abs=1e-2)


def test_ref_basic_apply_and_masking():
    """ reference tests/test_kernel_model.py:84-117, 206-242. """
    import cv2
    ra100, ra50 = _conftest_rasters()
    for kernel_shape, mask_partial in [((5, 5), False), ((1, 1), True), ((3, 3), True), ((3, 5), True), ((5, 5), True)]:
        km = RefSpaceModel(Model.gain_blk_offset, kernel_shape, mask_partial=mask_partial)
        param_ra = ra100.copy()
        mask = param_ra.mask
        param_ra.array = np.ones((2, *param_ra.shape), dtype='float32')
        param_ra.mask = mask
        out_ra = km.apply(ra50, param_ra)
        assert out_ra.shape == ra50.shape and out_ra.transform == ra50.transform
        if not mask_partial:
            assert (ra50.mask == out_ra.mask).all()
            assert out_ra.array[out_ra.mask] == pytest.approx(ra50.array[out_ra.mask] + 1, abs=1e-2)
        else:
            assert ra50.mask.sum() > out_ra.mask.sum() and ra50.mask[out_ra.mask].all()
            covered = (ra50.mask.reshape(20, 2, 10, 2).mean(axis=(1, 3)) >= 1).astype('uint8') & mask
            eroded = cv2.erode(covered, np.ones(np.add(kernel_shape, 2)), borderType=cv2.BORDER_CONSTANT, borderValue=0)
            assert (np.kron(eroded, np.ones((2, 2))).astype(bool) == out_ra.mask).all()


def _inpaint_scenario(kernel_shape):
    """ The rasters of reference tests/test_kernel_model.py:170-180: src == ref except one -100 pixel in the middle. """
    _, ra50 = _conftest_rasters()
    src_ra, ref_ra = ra50, ra50.copy()
    loc = np.floor(np.array(ref_ra.shape) / 2).astype(int)
    ul = (loc - np.floor(np.array(kernel_shape) / 2)).astype(int)
    low = np.zeros(ref_ra.shape, bool)
    low[ul[0]:ul[0] + kernel_shape[0], ul[1]:ul[1] + kernel_shape[1]] = True
    ref_ra.array[loc[0], loc[1]] = -100
    return src_ra, ref_ra, low


@pytest.mark.parametrize('kernel_shape', [(5, 5), (5, 7), (9, 9)])
def test_r2_inpainting(kernel_shape):
    """ reference tests/test_kernel_model.py:166-203, at the kernels the reference runs it with (on this 40 x 20 raster the
    scenario's own `R2 < .5 around the bad pixel` stops holding from (11, 11) on -- for the reference too). """
    src_ra, ref_ra, low = _inpaint_scenario(kernel_shape)
    no_inp = RefSpaceModel(Model.gain_offset, kernel_shape=kernel_shape, r2_inpaint_thresh=-np.inf).fit(src_ra, ref_ra)
    inp = RefSpaceModel(Model.gain_offset, kernel_shape=kernel_shape, r2_inpaint_thresh=0.5).fit(src_ra, ref_ra)
    for param_ra in (no_inp, inp):
        assert param_ra.array[2, ~low & ref_ra.mask] == pytest.approx(1, abs=1e-3)
        assert (param_ra.array[2, low] < .5).all()
    assert no_inp.array[1, no_inp.mask] != pytest.approx(0, abs=1e-1)
    assert inp.array[1, inp.mask] == pytest.approx(0, abs=1e-1)
    assert inp.array[0, inp.mask].var() < no_inp.array[0, no_inp.mask].var()


@pytest.mark.parametrize('kernel_shape', [(3, 3), (3, 5), (5, 5), (5, 7), (7, 7), (9, 9), (11, 11), (13, 13), (15, 15)])
@pytest.mark.parametrize('thresh', [0.5, 0.25])
def test_r2_inpainting_vs_oracle(kernel_shape, thresh):
    """ The same scenario for every kernel from (3, 3) to (15, 15), in-painted parameters against the oracle (the
    reference's own numpy / cv2 statements + the nodata-fill restatement): masks, infinities and NaNs in the same places,
    values within 1e-4.  (At (3, 3) corner windows of this raster hold a single distinct value: den = 0, R2 = NaN on both
    sides -- which is why the reference's own scenario starts at (5, 5).) """
    import warnings as _w
    kmnp = _oracle()
    src_ra, ref_ra, low = _inpaint_scenario(kernel_shape)
    with _w.catch_warnings():
        _w.simplefilter('ignore')
        got = RefSpaceModel(Model.gain_offset, kernel_shape=kernel_shape, r2_inpaint_thresh=thresh).fit(src_ra, ref_ra)
        exp = kmnp.refspace_fit(src_ra.array, tuple(src_ra.transform), float('nan'), ref_ra.array,
                                tuple(ref_ra.transform), float('nan'), Model.gain_offset, kernel_shape, False, thresh)
    got_a = got.array
    assert got_a.shape == exp.shape == (3,) + ref_ra.shape
    for b in range(3):
        assert np.array_equal(np.isnan(got_a[b]), np.isnan(exp[b])), f'band {b}: NaN pattern differs'
        assert np.array_equal(np.isinf(got_a[b]), np.isinf(exp[b])), f'band {b}: infinities differ'
    fin = np.isfinite(exp[0]) & np.isfinite(exp[1])
    assert rel_err(got_a[0][fin], exp[0][fin], 1e-3 * np.abs(exp[0][fin]).mean()) <= RTOL
    assert rel_err(got_a[1][fin], exp[1][fin], np.abs(exp[0][fin]) * float(np.nanmean(src_ra.array))) <= RTOL
    fin2 = np.isfinite(exp[2])
    assert np.max(np.abs(got_a[2][fin2] - exp[2][fin2])) <= 1e-4


def test_grid_mismatch_raises():
    ra100, ra50 = _conftest_rasters()
    km = KernelModel(Model.gain, (3, 3))
    with pytest.raises(ValueError):
        km.fit(ra50, ra100)
    with pytest.raises(ValueError):
        km.apply(ra50, ra100)


@pytest.mark.parametrize('model, kernel_shape, thresh', [
    (Model.gain_offset, (15, 15), 0.25), (Model.gain_blk_offset, (5, 5), None), (Model.gain, (1, 1), None),
])
def test_fused_call_equals_fit_then_apply(model, kernel_shape, thresh):
    """ RefSpaceModel.fuse (one hb_fuse_refspace call per band, what RasterFuse.process uses) runs the same kernels in
    the same order as fit() followed by apply(): bit-identical corrected pixels and parameters. """
    from homonim_b200 import RasterArray, RefSpaceModel
    src_ra, ref_ra = make_pair(110, 93, 20, bands=1, dtype='uint16', mu=3000.0, seed=9, device='cuda', src_nodata=0)
    src = RasterArray(src_ra.array[0].contiguous(), src_ra.crs, src_ra.transform, nodata=0)
    ref = RasterArray(ref_ra.array[0].contiguous(), ref_ra.crs, ref_ra.transform, nodata=float('nan'))
    km = RefSpaceModel(model, kernel_shape, find_r2=True, r2_inpaint_thresh=thresh)
    assert km.can_fuse(src, ref)
    params = km.fit(src, ref)
    corr = km.apply(src, params)
    corr_f, params_f = km.fuse(src, ref, want_params=True)
    assert torch.equal(corr_f.array.nan_to_num(-1.0), corr.array.nan_to_num(-1.0))
    assert torch.equal(torch.isnan(corr_f.array), torch.isnan(corr.array))
    assert torch.equal(params_f.array.nan_to_num(-7.0, posinf=-8.0, neginf=-9.0),
                       params.array.nan_to_num(-7.0, posinf=-8.0, neginf=-9.0))
    corr_n, none = km.fuse(src, ref)
    assert none is None and torch.equal(corr_n.array.nan_to_num(-1.0), corr.array.nan_to_num(-1.0))


def test_process_plan_cache_follows_in_place_edits():
    """ RasterFuse.process caches per-band prepared inputs between calls; in-place edits of the source or reference
    tensors (and a different band selection) must be picked up. """
    src_ra, ref_ra = make_pair(64, 56, 8, bands=2, dtype='uint16', mu=3000.0, seed=13, device='cuda', src_nodata=0)
    kw = dict(model=Model.gain_blk_offset, kernel_shape=(5, 5))
    with RasterFuse(src_ra, ref_ra) as fuse:
        first, _ = fuse.process(**kw)
        again, _ = fuse.process(**kw)
        assert torch.equal(first.array.nan_to_num(-1), again.array.nan_to_num(-1))
        ref_ra.array.mul_(1.5)                                   # in place: same storage, new version counter
        scaled, _ = fuse.process(**kw)
        fresh_src, fresh_ref = RasterArray(src_ra.array.clone(), src_ra.crs, src_ra.transform, nodata=0), \
            RasterArray(ref_ra.array.clone(), ref_ra.crs, ref_ra.transform, nodata=float('nan'))
        with RasterFuse(fresh_src, fresh_ref) as fuse2:
            expect, _ = fuse2.process(**kw)
        assert torch.equal(scaled.array.nan_to_num(-1), expect.array.nan_to_num(-1))
        assert not torch.equal(scaled.array.nan_to_num(-1), first.array.nan_to_num(-1))
        src_ra.array[0, :40, :40] = 0                            # new nodata in band 1 of the source
        holed, _ = fuse.process(**kw)
        assert bool(torch.isnan(holed.array[0, :40, :40]).all())
        assert torch.equal(holed.array[1].nan_to_num(-1), scaled.array[1].nan_to_num(-1))


def _random_pair(rng, dtype):
    """ A small source / reference pair on random (mis-aligned, non-integer ratio) north-up grids, numpy arrays. """
    from homonim_b200 import Affine
    ratio = float(rng.choice([2.0, 2.5, 3.0, 4.0, 5.0, 6.4, 8.0, 12.0, 20.0]))
    hp, wp = int(rng.integers(14, 30)), int(rng.integers(14, 30))
    ws = int(wp * ratio) // 4 * 4 if rng.random() < 0.7 else int(wp * ratio) - int(rng.integers(1, 4))
    hs = int(hp * ratio) - int(rng.integers(0, 3))
    mu = {'uint8': 120.0, 'uint16': 3000.0, 'float32': 0.3}[dtype]
    yy, xx = np.mgrid[0:hs, 0:ws]
    tex = np.sin(xx / (3.0 * ratio)) * np.cos(yy / (4.0 * ratio)) + 0.3 * np.sin((xx + 2 * yy) / (1.7 * ratio))
    src = mu * (1.0 + 0.3 * tex) + 0.02 * mu * rng.standard_normal((hs, ws))
    nodata = float('nan') if dtype == 'float32' else 0
    if dtype != 'float32':
        src = np.clip(np.round(src), 1, np.iinfo(dtype).max)
    src = src.astype(dtype)
    for _ in range(int(rng.integers(0, 4))):                       # nodata rectangles, some touching the border
        y, x = int(rng.integers(0, hs)), int(rng.integers(0, ws))
        src[y:y + int(rng.integers(2, 4 * ratio)), x:x + int(rng.integers(2, 4 * ratio))] = nodata
    # reference: 3 pixels larger than the source on every side, origin shifted by a random fraction of a source pixel
    res = 0.5
    ref_res = res * ratio
    shift = (float(rng.choice([0.0, 0.0, rng.random() * ratio])), float(rng.choice([0.0, 0.0, rng.random() * ratio])))
    src_tf = Affine(res, 0, 1000.0, 0, -res, 5000.0)
    ref_tf = Affine(ref_res, 0, 1000.0 - 3 * ref_res - shift[0] * res, 0, -ref_res, 5000.0 + 3 * ref_res + shift[1] * res)
    hr, wr = hp + 7, wp + 7
    ry, rx = np.mgrid[0:hr, 0:wr]
    gain = 0.6 + 0.2 * np.sin(rx / 5.0 + ry / 9.0)
    ref = (gain * mu * (1.0 + 0.3 * np.sin((rx - 3) / 3.0) * np.cos((ry - 3) / 4.0)) + 0.05 * mu
           + 0.01 * mu * rng.standard_normal((hr, wr))).astype('float32')
    return src, src_tf, nodata, ref, ref_tf


@pytest.mark.parametrize('seed', range(24))
def test_refspace_random_geometry_vs_oracle(seed):
    """ proc_crs = ref on random grids (non-integer ratios, sub-pixel offsets, widths that do / do not take the fast
    kernels, nodata touching the borders): masks exact, parameters and corrected pixels within 1e-4 of the oracle. """
    kmnp = _oracle()
    rng = np.random.default_rng(1000 + seed)
    dtype = ['uint8', 'uint16', 'float32'][seed % 3]
    model, kernel_shape, thresh = [(Model.gain, (1, 1), None), (Model.gain_blk_offset, (5, 5), None),
                                   (Model.gain_offset, (5, 5), 0.25), (Model.gain_blk_offset, (3, 7), None)][seed % 4]
    src, src_tf, nodata, ref, ref_tf = _random_pair(rng, dtype)
    crs = CRS.from_epsg(32735)
    src_ra = RasterArray(torch.from_numpy(src).cuda(), crs, src_tf, nodata=nodata)
    ref_ra = RasterArray(torch.from_numpy(ref).cuda(), crs, ref_tf, nodata=float('nan'))
    with RasterFuse(src_ra, ref_ra, proc_crs=ProcCrs.ref) as fuse:
        corr_ra, param_ra = fuse.process(model=model, kernel_shape=kernel_shape, param_filename='p',
                                         model_config=dict(r2_inpaint_thresh=thresh))
    with np.errstate(all='ignore'):
        exp_params, _, exp_corr = kmnp.fuse_band_blocks(src, tuple(src_tf), nodata, ref, tuple(ref_tf), float('nan'),
                                                        model, kernel_shape, 'ref', True, thresh)
    got_params = param_ra.to_host().array
    valid_src = src[src != nodata] if not np.isnan(nodata) else src[~np.isnan(src)]
    if kernel_shape == (1, 1):
        got_params, exp_params = got_params[:2], exp_params[:2]
    check_params(got_params, exp_params, float(np.mean(valid_src.astype('float64'))), f'seed {seed} params')
    check_corr(corr_ra.to_host().array[0] if corr_ra.array.ndim == 3 else corr_ra.to_host().array,
               exp_corr.astype('float32'), f'seed {seed} corr')


@pytest.mark.parametrize('model, kernel_shape, thresh', [
    (Model.gain_offset, (7, 5), None), (Model.gain_blk_offset, (5, 5), None), (Model.gain, (3, 3), None),
])
def test_srcspace_fused_equals_fit_then_apply(model, kernel_shape, thresh):
    """ SrcSpaceModel.fuse (apply fused into the fit kernel's epilogue, hb_fit_apply_same_grid -- what RasterFuse.process
    uses when no parameter raster is asked for) is bit-identical to fit() followed by apply(). """
    from homonim_b200 import SrcSpaceModel
    src_ra, ref_ra = make_pair(150, 131, 2, bands=1, dtype='float32', mu=0.3, seed=4, device='cuda',
                               src_nodata=float('nan'))
    src = RasterArray(src_ra.array[0].contiguous(), src_ra.crs, src_ra.transform, nodata=float('nan'))
    ref = RasterArray(ref_ra.array[0].contiguous(), ref_ra.crs, ref_ra.transform, nodata=float('nan'))
    km = SrcSpaceModel(model, kernel_shape, r2_inpaint_thresh=thresh)
    assert km.can_fuse(src, ref)
    corr = km.apply(src, km.fit(src, ref))
    fused = km.fuse(src, ref)
    assert torch.equal(torch.isnan(fused.array), torch.isnan(corr.array))
    assert torch.equal(fused.array.nan_to_num(-1.0, posinf=-2.0, neginf=-3.0),
                       corr.array.nan_to_num(-1.0, posinf=-2.0, neginf=-3.0))
    assert not SrcSpaceModel(Model.gain_offset, (5, 5), r2_inpaint_thresh=0.25).can_fuse(src, ref)
    # through RasterFuse.process: without a parameter raster the fused kernel runs, with one the two-step path does
    # (both on the reference block cropped to the source extent)
    with RasterFuse(src_ra, ref_ra, proc_crs=ProcCrs.src) as fuse:
        kw = dict(model=model, kernel_shape=kernel_shape, model_config=dict(r2_inpaint_thresh=thresh))
        got, none = fuse.process(**kw)
        two_step, params = fuse.process(param_filename='p', **kw)
    assert none is None and params is not None
    assert torch.equal(got.array.nan_to_num(-1.0), two_step.array.nan_to_num(-1.0))


def test_raster_compare_vs_reference_golden():
    """
    RasterCompare.process against the unmodified reference's RasterCompare.process (tests/golden/compare_stats.*,
    oracle/make_golden_compare.py) and against the oracle's sums.  Tolerances: statistics 1e-4 relative (the reference
    sums float32 terms in float32, the CUDA path in double); sums 1e-5 relative to the oracle's double sums of the same
    float32 terms (the re-projected planes agree to float32 rounding); identical pixel counts.
    """
    import json
    import pathlib
    from homonim_b200 import Affine, RasterCompare
    from homonim_b200.errors import IoError
    from oracle import kernel_model_np as knp
    NAN = float('nan')
    golden = pathlib.Path(__file__).resolve().parent / 'golden'
    meta = json.loads((golden / 'compare_stats.json').read_text())
    names = ['B4', 'B3', 'B2']
    crs = CRS0
    with np.load(golden / 'compare_stats.npz') as data:
        for ci in range(len(meta)):
            case = meta[f'case{ci}']
            src, ref = data[f'src{ci}'], data[f'ref{ci}']
            src_tf, ref_tf = Affine(*case['src_transform'][:6]), Affine(*case['ref_transform'][:6])
            for on_device in (True, False):
                to = (lambda a: torch.from_numpy(a).cuda()) if on_device else (lambda a: a)
                cmp = RasterCompare(RasterArray(to(src), crs, src_tf, nodata=NAN),
                                    RasterArray(to(ref), crs, ref_tf, nodata=NAN), proc_crs=case['proc_crs'],
                                    band_names=names)
                with pytest.raises(IoError):
                    cmp.process()
                with cmp:
                    stats = cmp.process()
                    sums = [cmp._band_sums_device(b).cpu().numpy() for b in range(3)]
                assert list(stats.keys()) == names + ['Mean']
                for band, band_stats in case['stats'].items():
                    assert stats[band]['n'] == int(band_stats['n']), (ci, band)
                    for key in ('r2', 'rmse', 'rrmse'):
                        assert abs(stats[band][key] - band_stats[key]) <= 1e-4 * abs(band_stats[key]), (ci, band, key)
                for b in range(3):
                    exp = knp.compare_band(src[b], tuple(src_tf), NAN, ref[b], tuple(ref_tf), NAN, case['proc_crs'],
                                           dtype='float64')
                    for key, value in zip(knp.COMPARE_SUM_KEYS, sums[b]):
                        assert abs(value - float(exp[key])) <= 1e-5 * abs(float(exp[key])), (ci, b, key)
            assert isinstance(RasterCompare.stats_table(stats), str)


def test_reference_published_table_gpu():
    """
    The CUDA path against known answers PUBLISHED by the reference (docs/cli.rst:58-72, real rasterio + GDAL pipeline):
    RasterFuse.process(gain-blk-offset, 5x5) of ngi_rgb_byte_1.tif with the Sentinel-2 reference, then RasterCompare of
    the source and of the corrected image with the Landsat-8 reference (tests/golden/docs_cli_ngi1.*,
    oracle/make_golden_docs.py).  Pixel counts must be identical; r2 / RMSE / rRMSE must agree with the three printed
    decimals to within half a unit of the last digit plus the 1e-4 relative float32 tolerance of the path.
    """
    import json
    import pathlib
    from homonim_b200 import Affine, RasterCompare
    golden = pathlib.Path(__file__).resolve().parent / 'golden'
    meta = json.loads((golden / 'docs_cli_ngi1.json').read_text())
    with np.load(golden / 'docs_cli_ngi1.npz') as data:
        src, s2, l8 = data['src'], data['s2'], data['l8']
    src_tf, s2_tf, l8_tf = (Affine(*meta[k][:6]) for k in ('src_transform', 's2_transform', 'l8_transform'))
    for on_device in (False, True):
        to = (lambda a: torch.from_numpy(a).cuda()) if on_device else (lambda a: a)
        src_ra = RasterArray(to(src), CRS0, src_tf, nodata=0)
        s2_ra = RasterArray(to(s2), CRS0, s2_tf, nodata=float('nan'))     # no nodata tag: the reference's reader uses NaN
        l8_ra = RasterArray(to(l8), CRS0, l8_tf, nodata=0)
        with RasterFuse(src_ra, s2_ra) as fuse:
            corr_ra, _ = fuse.process(model=Model.gain_blk_offset, kernel_shape=(5, 5))
        rows = {}
        for key, ra in (('source', src_ra), ('corrected', corr_ra)):
            with RasterCompare(ra, l8_ra, band_names=meta['band_names']) as cmp:
                rows[key] = cmp.process()['Mean']
        for key in ('source', 'corrected'):
            r2, rmse, rrmse, n = meta['published'][key]
            got = rows[key]
            assert got['n'] == n, (on_device, key, got)
            for name, exp in (('r2', r2), ('rmse', rmse), ('rrmse', rrmse)):
                assert abs(got[name] - exp) <= 0.5e-3 + 1e-4 * abs(exp), (on_device, key, name, got[name], exp)
            # and against the oracle's unrounded values of the same table: 1e-4 relative
            o_r2, o_rmse, o_rrmse, _ = meta['oracle'][key]
            for name, exp in (('r2', o_r2), ('rmse', o_rmse), ('rrmse', o_rrmse)):
                assert abs(got[name] - exp) <= 1e-4 * abs(exp), (on_device, key, name, got[name], exp)

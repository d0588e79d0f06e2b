"""
CPU tier: the oracle port (oracle/kernel_model_np.py) is pinned BIT-FOR-BIT against the golden vectors that
oracle/make_golden.py generated from the unmodified reference (/root/reference/homonim/kernel_model.py).
"""
import pathlib

import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import kernel_model_np as kmnp


def _same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize('name', golden_names('same'))
def test_same_grid_port_matches_reference(name):
    meta, g = load_golden(name)
    params = kmnp.fit_same_grid(g['src'], meta['src_nodata'], g['ref'], meta['ref_nodata'], meta['model'],
                                meta['kernel_shape'], meta['find_r2'], meta['r2_inpaint_thresh'])
    assert _same(params, g['params'])
    corr = kmnp.apply_same_grid(g['src'], params)
    assert _same(corr, g['corr'])


@pytest.mark.parametrize('name', golden_names('refspace'))
def test_refspace_port_matches_reference(name):
    meta, g = load_golden(name)
    params, param_tf, corr = kmnp.fuse_band_blocks(
        g['src'], meta['src_transform'], meta['src_nodata'], g['ref'], meta['ref_transform'], meta['ref_nodata'],
        meta['model'], meta['kernel_shape'], 'ref', meta['find_r2'], meta['r2_inpaint_thresh'], meta['mask_partial'])
    assert np.allclose(param_tf, meta['param_transform'])
    assert _same(params, g['params'])
    assert _same(corr.astype('float32'), g['corr'].astype('float32'))


@pytest.mark.parametrize('name', golden_names('srcspace'))
def test_srcspace_port_matches_reference(name):
    meta, g = load_golden(name)
    params, _, corr = kmnp.fuse_band_blocks(
        g['src'], meta['src_transform'], meta['src_nodata'], g['ref'], meta['ref_transform'], meta['ref_nodata'],
        meta['model'], meta['kernel_shape'], 'src', meta['find_r2'], meta['r2_inpaint_thresh'], meta['mask_partial'])
    assert _same(params, g['params'])
    assert _same(corr, g['corr'])


def test_golden_fixture_inventory():
    # every reference model, R2 on/off, in-painting, all dtypes and both processing grids are pinned
    names = golden_names()
    assert len(names) >= 26
    for token in ('gain_k', 'gain-blk-offset', 'gain-offset', '_r2', '_inp', 'uint8', 'uint16', 'float32', 'partial',
                  'refspace', 'srcspace', 'conftest'):
        assert any(token in n for n in names), token


def test_oracle_convert_dtype_matches_reference():
    """ oracle.kernel_model_np.convert_dtype against RasterArray._convert_array_dtype of the unmodified reference
    (tests/golden/convert_dtype.npz, oracle/make_golden_convert.py). """
    import pathlib
    from oracle import kernel_model_np as kmnp
    with np.load(pathlib.Path(__file__).resolve().parent / 'golden' / 'convert_dtype.npz') as data:
        corr = data['corr']
        for key in data.files:
            if key == 'corr':
                continue
            dtype, nodata = key.rsplit('_', 1)
            got = kmnp.convert_dtype(corr, dtype, float(nodata))
            assert got.dtype == data[key].dtype
            assert np.array_equal(got, data[key]), key


def test_oracle_compare_matches_reference():
    """ The oracle's RasterCompare restatement (compare.py:142-187, 232-256) against the unmodified reference's
    RasterCompare.process (tests/golden/compare_stats.*): bit-exact with float32 summation (what the reference does);
    within 1e-5 when the same float32 terms are summed in double (what the CUDA path does). """
    import json
    import pathlib
    GOLDEN_DIR = pathlib.Path(__file__).resolve().parent / 'golden'
    meta = json.loads((GOLDEN_DIR / 'compare_stats.json').read_text())
    names = ['B4', 'B3', 'B2']
    with np.load(GOLDEN_DIR / 'compare_stats.npz') as data:
        for ci in range(len(meta)):
            case = meta[f'case{ci}']
            src, ref = data[f'src{ci}'], data[f'ref{ci}']
            args = (tuple(case['src_transform']), float('nan'))
            for dtype, tol in (('float32', 0.0), ('float64', 1e-5)):
                sums = [kmnp.compare_band(src[b], args[0], args[1], ref[b], tuple(case['ref_transform']), float('nan'),
                                         case['proc_crs'], dtype=dtype) for b in range(3)]
                stats = kmnp.compare_image_stats(sums, names)
                assert list(stats.keys()) == names + ['Mean']
                for band, band_stats in case['stats'].items():
                    for key, value in band_stats.items():
                        assert abs(stats[band][key] - value) <= tol * abs(value), (ci, dtype, band, key)
                if dtype == 'float32':
                    for b in range(3):
                        for key, value in case['image_sums'][b].items():
                            assert float(sums[b][key]) == value, (ci, b, key)


def _docs_rows(kmnp_mod, src, src_tf, s2, s2_tf, l8, l8_tf, names):
    """ (source, corrected) rows of the docs/cli.rst table for one image, computed with the oracle. """
    nan = float('nan')
    src_sums, corr_sums = [], []
    for b in range(3):
        src_sums.append(kmnp_mod.compare_band(src[b], src_tf, 0.0, l8[b], l8_tf, 0.0, 'ref'))
        _, _, corr = kmnp_mod.fuse_band_blocks(src[b], src_tf, 0.0, s2[b], s2_tf, nan, 'gain-blk-offset', (5, 5))
        corr_sums.append(kmnp_mod.compare_band(corr, src_tf, nan, l8[b], l8_tf, 0.0, 'ref'))
    return {key: kmnp_mod.compare_image_stats(sums, names)['Mean']
            for key, sums in (('source', src_sums), ('corrected', corr_sums))}


def test_oracle_reproduces_reference_published_table():
    """
    Known answers published by the reference (docs/cli.rst:58-72): `homonim fuse -m gain-blk-offset -k 5 5` of
    ngi_rgb_byte_1.tif with the Sentinel-2 reference, then `homonim compare` of the source and the corrected image with
    the Landsat-8 reference -- numbers from the real rasterio + GDAL pipeline.  The oracle chain (GRA_Average
    restatement -> fit -> GRA_CubicSpline restatement -> apply -> compare sums) must reproduce every printed digit and
    the pixel count exactly.  This is what pins oracle/gdal_restate.c to real GDAL output.
    """
    import json
    import pathlib
    golden = pathlib.Path(__file__).resolve().parent / 'golden'
    meta = json.loads((golden / 'docs_cli_ngi1.json').read_text())
    with np.load(golden / 'docs_cli_ngi1.npz') as data:
        rows = _docs_rows(kmnp, data['src'], tuple(meta['src_transform']), data['s2'], tuple(meta['s2_transform']),
                          data['l8'], tuple(meta['l8_transform']), meta['band_names'])
    for key in ('source', 'corrected'):
        r2, rmse, rrmse, n = meta['published'][key]
        got = rows[key]
        assert int(got['n']) == n, key
        assert (round(float(got['r2']), 3), round(float(got['rmse']), 3), round(float(got['rrmse']), 3)) == \
            (r2, rmse, rrmse), (key, got)


@pytest.mark.skipif(not pathlib.Path('/root/reference/tests/data/source/ngi_rgb_byte_4.tif').exists(),
                    reason='container-only: reads the reference test images from /root/reference')
def test_oracle_reproduces_reference_published_table_all_images():
    """ The same for all four source images of docs/cli.rst:63-72, read from the reference checkout (32 published
    numbers). """
    from oracle import make_golden_docs as mgd
    (s2, s2_tf), (l8, l8_tf) = mgd.matched_references()
    for name, published in mgd.PUBLISHED.items():
        src = mgd.read_geotiff(mgd.DATA / 'source' / name)
        mgd.check_rows(mgd.docs_rows(src['array'], src['transform'], s2, s2_tf, l8, l8_tf), published, name)


# ---------------------------------------------------------------------------------------------------------------------
# GDALFillNodata restatement (in-painting, kernel_model.py:366): GDAL is not installable here, so parity with GDAL itself
# stays unpinned; what CAN be pinned is that the C restatement does what the algorithm's description says.
# ---------------------------------------------------------------------------------------------------------------------
def test_fillnodata_restatement_vs_independent_restatement():
    """ oracle/gdal_restate.c::gr_fillnodata against oracle/fillnodata_alt.py (brute force, written from the algorithm's
    description): bit-identical on random images / masks / search radii. """
    from oracle import gdal_restate as gr
    from oracle.fillnodata_alt import fillnodata_alt
    rng = np.random.default_rng(0)
    for case in range(16):
        h, w = int(rng.integers(4, 50)), int(rng.integers(4, 60))
        img = rng.normal(10, 3, (h, w)).astype('float32')
        mask = rng.random((h, w)) < rng.choice([0.01, 0.05, 0.3, 0.9])
        radius = float(rng.choice([2, 7, 100]))
        a, b = gr.fillnodata(img, mask, radius), fillnodata_alt(img, mask, radius)
        assert np.array_equal(a, b), f'case {case}: {int((a != b).sum())} pixels differ'


def test_fillnodata_properties():
    """ Properties of the four-quadrant inverse-distance fill that the description implies. """
    from oracle import gdal_restate as gr
    nan = np.float32('nan')
    # (1) no source at all / none within reach: untouched
    img = np.full((9, 9), 7.0, 'float32')
    assert np.array_equal(gr.fillnodata(img, np.zeros((9, 9), bool), 100.0), img)
    # (2) a single source at distance 1: the filled value IS the source value
    img = np.zeros((5, 5), 'float32'); mask = np.zeros((5, 5), bool)
    img[2, 1], mask[2, 1] = 3.25, True
    assert gr.fillnodata(img, mask, 100.0)[2, 2] == np.float32(3.25)
    # (3) the cut-off: a source exactly max_search_distance away fills, one pixel further does not
    img = np.zeros((1, 130), 'float32'); mask = np.zeros((1, 130), bool)
    img[0, 0], mask[0, 0] = 5.0, True
    out = gr.fillnodata(img, mask, 100.0)
    assert out[0, 100] == np.float32(5.0) and out[0, 101] == 0.0
    # (4) a source on the pixel's own ROW is found by the top and the bottom quadrant of its side: it counts twice;
    #     a source in its own COLUMN counts once (the centre column belongs to the left side only)
    img = np.zeros((7, 7), 'float32'); mask = np.zeros((7, 7), bool)
    img[3, 5], mask[3, 5] = 100.0, True         # same row, 2 to the right: top-right AND bottom-right quadrant
    img[1, 3], mask[1, 3] = 40.0, True          # same column, 2 above: top-left quadrant only
    got = gr.fillnodata(img, mask, 100.0)[3, 3]
    assert got == np.float32((40.0 / 2 + 2 * 100.0 / 2) / (3 / 2))           # (40 + 100 + 100) / 3 = 80
    # (4b) one source per quadrant at most, the nearest; at equal distance the one found at the smaller column step
    img = np.zeros((7, 7), 'float32'); mask = np.zeros((7, 7), bool)
    img[3, 1], mask[3, 1] = 10.0, True          # same row, 2 to the left (column step 2)
    img[1, 3], mask[1, 3] = 40.0, True          # same column, 2 above (column step 0): takes the top-left quadrant
    assert gr.fillnodata(img, mask, 100.0)[3, 3] == np.float32(25.0)        # (40 [tl] + 10 [bl]) / 2
    # (5) up-down mirror symmetry (up to the order of the four-term sum).  Left-right mirroring does NOT hold in general:
    #     the centre column belongs to the left side, so a source straight above / below changes sides -- also asserted
    rng = np.random.default_rng(4)
    img = rng.normal(5, 1, (23, 31)).astype('float32')
    mask = rng.random((23, 31)) < 0.08
    base = gr.fillnodata(img, mask, 100.0)
    flipped = np.flipud(gr.fillnodata(np.ascontiguousarray(np.flipud(img)), np.ascontiguousarray(np.flipud(mask)), 100.0))
    assert np.allclose(flipped, base, rtol=2e-6, atol=0)
    mirrored = np.fliplr(gr.fillnodata(np.ascontiguousarray(np.fliplr(img)), np.ascontiguousarray(np.fliplr(mask)), 100.0))
    differs = ~np.isclose(mirrored, base, rtol=2e-6, atol=0)
    assert differs.any() and not differs[:, ~mask.any(axis=0)].any()       # only in columns that hold a source
    # (6) sources are never modified, and only ORIGINAL sources are used (a filled pixel does not become a source)
    out = gr.fillnodata(img, mask, 3.0)
    assert np.array_equal(out[mask], img[mask])
    far = np.ones(mask.shape, bool)
    ys, xs = np.nonzero(mask)
    for y, x in zip(ys, xs):
        far[max(y - 3, 0):y + 4, max(x - 3, 0):x + 4] = False
    assert np.array_equal(out[far], img[far])   # nothing within reach (Chebyshev > 3 implies Euclidean > 3): untouched
    assert not np.isnan(out).any() and nan != nan

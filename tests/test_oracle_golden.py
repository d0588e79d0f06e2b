"""
CPU tier: the oracle port (oracle/kernel_model_np.py) is pinned BIT-FOR-BIT against the golden vectors that
oracle/make_golden.py generated from the unmodified reference (/root/reference/homonim/kernel_model.py).
"""
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

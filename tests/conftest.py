"""
Test configuration.  Tiers:
  * ``-m "not gpu"``: oracle pinned against tests/golden (generated from the unmodified reference), host logic, C-ABI
    symbol checks, world_size-2 gloo tests.  No CUDA device is needed (and none is used).
  * ``-m gpu``: parity of the CUDA path (through the Python API and the C-ABI) against the golden fixtures and the
    oracle port, plus full-size property checks.
"""
import json
import pathlib
import sys

import numpy as np
import pytest

REPO = pathlib.Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN_DIR = REPO / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _decode_meta(meta):
    meta = dict(meta)
    for key in ('src_nodata', 'ref_nodata'):
        if meta.get(key) == 'nan':
            meta[key] = float('nan')
    if meta.get('r2_inpaint_thresh') == '-inf':
        meta['r2_inpaint_thresh'] = float('-inf')
    meta['kernel_shape'] = tuple(meta['kernel_shape'])
    return meta


def golden_index():
    return {k: _decode_meta(v) for k, v in json.loads((GOLDEN_DIR / 'index.json').read_text()).items()}


def golden_names(kind=None):
    return sorted(k for k, v in golden_index().items() if kind is None or v['kind'] == kind)


def load_golden(name):
    meta = golden_index()[name]
    with np.load(GOLDEN_DIR / f'{name}.npz') as data:
        arrays = {k: data[k] for k in data.files}
    return meta, arrays


def assert_same_mask(actual, expected, what=''):
    a, e = np.isnan(actual), np.isnan(expected)
    assert a.shape == e.shape, f'{what}: shape {a.shape} != {e.shape}'
    assert np.array_equal(a, e), f'{what}: nodata masks differ at {int((a != e).sum())} pixels'


def rel_err(actual, expected, floor):
    """ max |a - e| / max(|e|, floor) over finite expected values; infinities must match exactly. """
    actual, expected = np.asarray(actual, dtype='float64'), np.asarray(expected, dtype='float64')
    inf = np.isinf(expected)
    assert np.array_equal(actual[inf], expected[inf]), 'infinite values differ'
    ok = np.isfinite(expected)
    assert np.all(np.isfinite(actual[ok])), 'non-finite result where the reference is finite'
    if not ok.any():
        return 0.0
    floor = np.broadcast_to(np.asarray(floor, dtype='float64'), expected.shape)
    return float(np.max(np.abs(actual[ok] - expected[ok]) / np.maximum(np.abs(expected[ok]), floor[ok])))


# float32 tolerance of BASELINE.json's north_star: param and corrected-pixel max rel err <= 1e-4, masks identical
RTOL = 1e-4


def check_params(actual, expected, src_mean, what=''):
    """
    Parity metric of SURVEY.md 8(d): masks exact per band; gain / corrected relative error with a floor of 1e-3 x the
    band mean; offset error relative to max(|offset|, |gain| * mean(src)) (the offset is a small difference of large
    terms, so a purely relative test on it is ill-posed); R2 absolute.
    """
    assert actual.shape == expected.shape, f'{what}: shape {actual.shape} != {expected.shape}'
    for b in range(expected.shape[0]):
        assert_same_mask(actual[b], expected[b], f'{what} band {b}')
    gain_e = expected[0]
    g_floor = 1e-3 * np.nanmean(np.abs(gain_e[np.isfinite(gain_e)])) if np.isfinite(gain_e).any() else 1.0
    assert rel_err(actual[0], gain_e, g_floor) <= RTOL, f'{what}: gain'
    o_floor = np.maximum(np.abs(np.nan_to_num(gain_e, nan=0.0, posinf=0.0, neginf=0.0)) * abs(src_mean), 1e-30)
    assert rel_err(actual[1], expected[1], o_floor) <= RTOL, f'{what}: offset'
    if expected.shape[0] > 2:
        fin = np.isfinite(expected[2])
        assert np.array_equal(actual[2][~fin & ~np.isnan(expected[2])], expected[2][~fin & ~np.isnan(expected[2])])
        if fin.any():
            assert np.max(np.abs(actual[2][fin].astype('float64') - expected[2][fin])) <= RTOL, f'{what}: R2'


def check_corr(actual, expected, what=''):
    assert_same_mask(actual, expected, what)
    fin = np.isfinite(expected)
    floor = 1e-3 * np.mean(np.abs(expected[fin])) if fin.any() else 1.0
    assert rel_err(actual, expected, floor) <= RTOL, f'{what}: corrected pixels'

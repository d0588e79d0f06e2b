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


# float32 tolerance of BASELINE.json's north_star: param and corrected-pixel max rel err <= 1e-4, masks identical
RTOL = 1e-4


def assert_same_mask(actual, expected, what=''):
    a, e = np.isnan(actual), np.isnan(expected)
    assert a.shape == e.shape, f'{what}: shape {a.shape} != {e.shape}'
    assert np.array_equal(a, e), f'{what}: nodata masks differ at {int((a != e).sum())} pixels'


def rel_err_map(actual, expected, floor):
    """ |a - e| / max(|e|, floor) per pixel (0 where the reference is nan); infinities must match exactly. """
    actual, expected = np.asarray(actual, dtype='float64'), np.asarray(expected, dtype='float64')
    inf = np.isinf(expected)
    assert np.array_equal(actual[inf], expected[inf]), 'infinite values differ'
    ok = np.isfinite(expected)
    assert np.all(np.isfinite(actual[ok])), 'non-finite result where the reference is finite'
    floor = np.broadcast_to(np.asarray(floor, dtype='float64'), expected.shape)
    err = np.zeros(expected.shape)
    err[ok] = np.abs(actual[ok] - expected[ok]) / np.maximum(np.abs(expected[ok]), floor[ok])
    return err


def rel_err(actual, expected, floor):
    """ max |a - e| / max(|e|, floor) over finite expected values; infinities must match exactly. """
    err = rel_err_map(actual, expected, floor)
    return float(err.max()) if err.size else 0.0


def ill_conditioned(gain_expected):
    """
    Pixels where the reference's own solve is ill-conditioned: a window whose denominator (sum of the normalised
    source, or N*sum(s^2) - sum(s)^2) crosses zero makes the reference return an arbitrarily large gain whose value is
    rounding noise.  They are recognised by |gain| > 50 x the band's median |gain|.
    """
    fin = np.isfinite(gain_expected)
    med = np.median(np.abs(gain_expected[fin])) if fin.any() else 1.0
    with np.errstate(invalid='ignore'):
        return np.abs(gain_expected) > 50 * max(med, 1e-30)


def _assert_within(err, tol, bad_ok, what, max_fraction=1e-4):
    """ Every pixel is within ``tol`` except at most a negligible number of ill-conditioned ones (``bad_ok``). """
    failing = err > tol
    if failing.any():
        assert bad_ok[failing].all(), f'{what}: max err {err[failing & ~bad_ok].max():.3g} > {tol:g} on ' \
                                      f'{int((failing & ~bad_ok).sum())} well-conditioned pixels'
        assert failing.sum() <= max(2, max_fraction * err.size), f'{what}: {int(failing.sum())} ill-conditioned pixels'


def check_params(actual, expected, src_mean, what=''):
    """
    Parity metric of SURVEY.md 8(d): masks exact per band; gain relative error with a floor of 1e-3 x the band mean;
    offset error relative to max(|offset|, |gain| * mean(src)) (the offset is a small difference of large terms, so a
    purely relative test on it is ill-posed); R2 absolute.  Tolerance 1e-4 throughout (BASELINE.json north_star).
    Pixels that miss it must be ill-conditioned in the reference itself and a negligible fraction.

    """
    assert actual.shape == expected.shape, f'{what}: shape {actual.shape} != {expected.shape}'
    for b in range(expected.shape[0]):
        assert_same_mask(actual[b], expected[b], f'{what} band {b}')
    gain_e = expected[0]
    bad = ill_conditioned(gain_e)
    fin = np.isfinite(gain_e)
    g_floor = 1e-3 * np.mean(np.abs(gain_e[fin & ~bad])) if (fin & ~bad).any() else 1.0
    _assert_within(rel_err_map(actual[0], gain_e, g_floor), RTOL, bad, f'{what}: gain')
    o_floor = np.maximum(np.abs(np.nan_to_num(gain_e, nan=0.0, posinf=0.0, neginf=0.0)) * abs(src_mean), 1e-30)
    _assert_within(rel_err_map(actual[1], expected[1], o_floor), RTOL, bad, f'{what}: offset')
    if expected.shape[0] > 2:
        r2_e = expected[2]
        nonfin = ~np.isfinite(r2_e) & ~np.isnan(r2_e)
        assert np.array_equal(actual[2][nonfin], r2_e[nonfin]), f'{what}: infinite R2 values differ'
        fin2 = np.isfinite(r2_e)
        err = np.zeros(r2_e.shape)
        err[fin2] = np.abs(actual[2][fin2].astype('float64') - r2_e[fin2])
        _assert_within(err, RTOL, bad, f'{what}: R2')
    return bad


def check_corr(actual, expected, what='', bad=None):
    """ Corrected pixels: masks exact, max rel err <= 1e-4 (floor 1e-3 x band mean); ``bad`` = ill-conditioned. """
    assert_same_mask(actual, expected, what)
    bad = np.zeros(expected.shape, bool) if bad is None else bad
    fin = np.isfinite(expected) & ~bad
    floor = 1e-3 * np.mean(np.abs(expected[fin])) if fin.any() else 1.0
    _assert_within(rel_err_map(actual, expected, floor), RTOL, bad, f'{what}: corrected pixels')

"""
CPU tier: host-side API surface (mirrors the reference's tests/test_kernel_model.py:296-334 and the RasterArray
semantics of tests/test_raster_array.py that the hot path relies on), and the C-ABI library: it must load and export
every symbol that include/homonim_b200.h declares.  No compute call is made here (there is no GPU).
"""
import ctypes
import re
import warnings

import numpy as np
import pytest

from conftest import REPO
from homonim_b200 import (Affine, CRS, KernelModel, Model, NativeLibraryError, ProcCrs, RasterArray, RasterFuse,
                          RefSpaceModel, Resampling, SrcSpaceModel, _native)
from homonim_b200.errors import ConfigWarning, ImageProfileError, IoError
from homonim_b200.fuse import ref_window_for_src
from homonim_b200.geometry import grid_map
from homonim_b200.kernel_model import overlap_for_kernel, validate_kernel_shape

CRS0 = CRS.from_epsg(3857)


def test_enums_match_reference_strings():
    assert [m.value for m in Model] == ['gain', 'gain-blk-offset', 'gain-offset']
    assert [p.value for p in ProcCrs] == ['auto', 'src', 'ref']
    assert Model('gain-offset') is Model.gain_offset and ProcCrs('ref') is ProcCrs.ref
    assert Resampling.coerce('cubic_spline') is Resampling.cubic_spline


@pytest.mark.parametrize('model, kernel_shape', [
    (Model.gain, (0, 0)), (Model.gain_blk_offset, (0, 0)), (Model.gain_offset, (4, 5)), (Model.gain_offset, (1, 1)),
    (Model.gain, (2, 3)),
])
def test_kernel_shape_exception(model, kernel_shape):
    with pytest.raises(ValueError):
        RefSpaceModel(model=model, kernel_shape=kernel_shape)


def test_kernel_shape_warning_and_overlap():
    with pytest.warns(ConfigWarning):
        assert validate_kernel_shape((3, 3), Model.gain_offset) == (3, 3)
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        assert validate_kernel_shape((5, 5), Model.gain_offset) == (5, 5)
    assert overlap_for_kernel((5, 5)) == (3, 3) and overlap_for_kernel((1, 15)) == (1, 8)


def test_config():
    config = dict(r2_inpaint_thresh=0.1, mask_partial=True, downsampling='bilinear', upsampling='nearest')
    km = RefSpaceModel(Model.gain, (5, 5), find_r2=True, **config)
    for key, value in config.items():
        assert getattr(km, '_' + key) == value
    assert km.model is Model.gain and km.kernel_shape == (5, 5) and km.find_r2 is True
    assert KernelModel.create_config() == dict(r2_inpaint_thresh=0.25, mask_partial=False,
                                               downsampling=Resampling.average, upsampling=Resampling.cubic_spline)
    assert KernelModel.default_model is Model.gain_blk_offset and KernelModel.default_kernel_shape == (5, 5)


def test_config_exception():
    with pytest.raises(TypeError):
        RefSpaceModel(Model.gain, (5, 5), find_r2=True, unknown='value')


def test_raster_array_semantics():
    a = np.arange(1, 13, dtype='float32').reshape(3, 4)
    a[0, 0] = np.nan
    ra = RasterArray(a.copy(), CRS0, Affine(2, 0, 10, 0, -2, 50))
    assert ra.shape == (3, 4) and ra.count == 1 and ra.dtype == 'float32' and ra.res == (2.0, 2.0)
    assert ra.bounds == (10.0, 44.0, 18.0, 50.0)
    assert ra.mask.sum() == 11 and not ra.mask[0, 0]
    ra.nodata = 5.0                                  # re-labelling nodata rewrites the masked pixels
    assert ra.array[0, 0] == 5.0 and ra.mask.sum() == 10
    prof = ra.profile
    assert prof['count'] == 1 and prof['width'] == 4 and prof['height'] == 3 and prof['nodata'] == 5.0
    ra2 = RasterArray.from_profile(None, dict(prof, count=2))
    assert ra2.array.shape == (2, 3, 4) and (ra2.array == 5.0).all() and not ra2.mask.any()
    with pytest.raises(ImageProfileError):
        RasterArray.from_profile(None, dict(crs=CRS0))
    with pytest.raises(ValueError):
        RasterArray(np.zeros(3), CRS0, Affine.identity())
    with pytest.raises(ValueError):
        ra.array = np.zeros((2, 2), dtype='float32')
    m = ra.mask_ra
    assert m.dtype == 'uint8' and m.nodata is None and m.mask.all()
    c = ra.copy()
    c.array[1, 1] = -1
    assert ra.array[1, 1] != -1
    ra3 = RasterArray(np.stack([a, a]), CRS0, Affine.identity())
    mask = ra3.mask.copy()
    mask[2, :] = False
    ra3.mask = mask
    assert np.isnan(ra3.array[:, 2, :]).all()


def test_geometry_grid_map_and_windows():
    tf100 = Affine(1, 0, 0, 0, -1, 0) * Affine.translation(5, 5)
    tf50 = tf100 * Affine.scale(0.5)
    assert grid_map(tf50, tf100) == (2.0, 0.0, 2.0, 0.0)
    assert grid_map(tf100, tf50)[:1] == (0.5,)
    assert (~tf50) * (tf50 * (3, 4)) == pytest.approx((3, 4))
    src = RasterArray(np.zeros((40, 20), 'float32'), CRS0, tf50 * Affine.translation(3, 5))
    ref = RasterArray(np.zeros((30, 30), 'float32'), CRS0, tf100 * Affine.translation(-4, -4))
    # source covers columns 1.5 .. 11.5 and rows 2.5 .. 22.5 of the tf100 grid -> ref pixels [5, 16) x [6, 27)
    assert ref_window_for_src(src, ref) == (6, 5, 21, 11)
    with pytest.raises(NotImplementedError):
        grid_map(Affine(1, 0.1, 0, 0, -1, 0), tf100)


def test_raster_fuse_host_logic():
    src = RasterArray(np.ones((2, 40, 20), 'float32'), CRS0, Affine(0.5, 0, 5, 0, -0.5, -5))
    ref = RasterArray(np.ones((3, 30, 30), 'float32'), CRS0, Affine(1, 0, 1, 0, -1, -1))
    fuse = RasterFuse(src, ref)
    assert fuse.proc_crs is ProcCrs.ref and fuse.src_bands == (1, 2) and fuse.ref_bands == (1, 2)
    assert RasterFuse(ref, src, src_bands=[1, 2]).proc_crs is ProcCrs.src
    with pytest.warns(ConfigWarning):
        RasterFuse(src, ref, proc_crs='src')
    with pytest.raises(IoError):
        fuse.process()
    with pytest.raises(FileNotFoundError):          # file names are opened as GeoTIFFs (tests/test_files_cpu.py)
        RasterFuse('src.tif', 'ref.tif')
    with pytest.raises(TypeError):
        RasterFuse(src, 'ref.tif')
    assert RasterFuse.create_block_config(threads=1) == dict(threads=1, max_block_mem=100)
    assert RasterFuse.create_out_profile()['dtype'] == 'float32'
    assert RasterFuse.create_model_config() == KernelModel.create_config()
    blk = fuse._ref_block(0)
    assert blk.shape == (20, 10) and blk.transform == Affine(1, 0, 5, 0, -1, -5)


# ---- the C-ABI library -----------------------------------------------------------------------------------------------
def _header_symbols():
    text = (REPO / 'include' / 'homonim_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(hb_[a-z0-9_]+)\s*\(', text)))


def test_capi_library_loads_and_exports_every_header_symbol():
    lib = _native.lib()
    symbols = _header_symbols()
    assert len(symbols) >= 16
    assert sorted(_native.SIGNATURES) == symbols, 'ctypes binding and header disagree'
    raw = ctypes.CDLL(str(_native.LIB_PATH))
    for name in symbols:
        assert hasattr(raw, name), f'{name} is declared in the header but not exported'
    assert lib.hb_abi_version() == _native.ABI_VERSION


def test_no_cpu_fallback():
    """ Without a CUDA device the product path must fail loudly, not compute on the CPU. """
    lib = _native.lib()
    if lib.hb_device_count() > 0:
        pytest.skip('a CUDA device is present')
    a = np.ones((8, 8), 'float32')
    ra = RasterArray(a, CRS0, Affine.identity())
    with pytest.raises(NativeLibraryError):
        KernelModel(Model.gain, (1, 1)).fit(ra, ra.copy())
    with pytest.raises(NativeLibraryError):
        SrcSpaceModel(Model.gain, (1, 1)).apply(ra, RasterArray(np.ones((2, 8, 8), 'float32'), CRS0, Affine.identity()))
    # a compute entry point called directly also refuses (bad arguments are rejected before any launch)
    assert lib.hb_downsample_average(None, 2, 8, 8, 0, 0.0, None, 4, 4, 2.0, 0.0, 2.0, 0.0, None) != 0
    assert b'bad arguments' in lib.hb_last_error()


def test_product_does_not_import_oracle():
    """ The oracle is test infrastructure: nothing under homonim_b200/ may reference it. """
    for path in (REPO / 'homonim_b200').rglob('*'):
        if path.suffix in ('.py', '.cu', '.cuh', '.h'):
            text = path.read_text()
            assert 'import oracle' not in text and 'from oracle' not in text, path


def test_bench_reference_arm_contract():
    """ `bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the contract's
    keys; run here on the tiny workload. """
    import json
    import pathlib
    import subprocess
    import sys
    repo = pathlib.Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(repo / 'bench.py'), '--impl', 'reference', '--workload', 'tiny', '--steps',
                          '2', '--warmup', '1'], capture_output=True, text=True, timeout=600, cwd=str(repo))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['unit'] == 'Mpix/s' and line['value'] > 0
    assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
    assert 'one_block' in line['cpu_baseline']['modes']
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in line['config'] and 'model' not in line['config']


def test_bench_reference_arm_on_real_images_bounded():
    """ BASELINE.json configs[0] (`--workload c1`: the reference's own test images, committed as a fixture) through the
    reference arm, with a time budget small enough to force the bounded-sample path (fewer bands per step). """
    import json
    import os
    import pathlib
    import subprocess
    import sys
    repo = pathlib.Path(__file__).resolve().parent.parent
    env = dict(os.environ, HB_REFERENCE_BUDGET_S='0.2')
    out = subprocess.run([sys.executable, str(repo / 'bench.py'), '--impl', 'reference', '--workload', 'c1', '--steps',
                          '2', '--warmup', '1'], capture_output=True, text=True, timeout=600, cwd=str(repo), env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith('{')][0])
    assert line['data'].startswith('real') and line['config']['src_shape'] == [1421, 805]
    assert line['config']['kernel_model'] == 'gain-blk-offset' and line['config']['bands'] == 3
    assert 'bounded sample' in line['cpu_baseline']['sample'] and line['value'] > 0
    import bench
    assert bench._small(bench.WORKLOADS['c1']) and bench._small(bench.WORKLOADS['tiny'])
    assert not bench._small(bench.WORKLOADS['c2']) and not bench._small(bench.WORKLOADS['c3'])
    assert 'flush' in bench._config(bench.WORKLOADS['c1'])['l2'] or 'written between' in bench._config(bench.WORKLOADS['c1'])['l2']


def test_raster_compare_host_logic():
    """ RasterCompare's host side (no GPU): configuration defaults, the closed-pair error, the statistics formed from
    the sums (same expressions as compare.py:145-187, checked against the oracle restatement) and the table. """
    import numpy as np
    from homonim_b200 import Affine, CRS, ProcCrs, RasterArray, RasterCompare, Resampling
    from oracle import kernel_model_np as knp
    config = RasterCompare.create_config()
    assert config['max_block_mem'] == 512 and config['downsampling'] == Resampling.average
    assert config['upsampling'] == Resampling.cubic_spline and config['threads'] >= 1
    crs = CRS.from_epsg(3857)
    src = RasterArray(np.ones((2, 40, 40), 'float32'), crs, Affine(5, 0, 0, 0, -5, 0))
    ref = RasterArray(np.ones((3, 20, 20), 'float32'), crs, Affine(10, 0, 0, 0, -10, 0))
    cmp = RasterCompare(src, ref)
    assert cmp.proc_crs == ProcCrs.ref and tuple(cmp.src_bands) == (1, 2) and tuple(cmp.ref_bands) == (1, 2)
    assert cmp._get_resampling(src.res, ref.res) == Resampling.average
    assert cmp._get_resampling(ref.res, src.res, upsampling='nearest') == Resampling.nearest
    with pytest.raises(IoError):
        cmp.process()
    rng = np.random.default_rng(3)
    image_sums = []
    for _ in range(2):
        a, b = rng.normal(5, 1, 500), rng.normal(6, 1, 500)
        image_sums.append(dict(src_sum=a.sum(), ref_sum=b.sum(), src2_sum=(a * a).sum(), ref2_sum=(b * b).sum(),
                               src_ref_sum=(a * b).sum(), res2_sum=((b - a) ** 2).sum(), mask_sum=500))
    stats = cmp._get_image_stats(image_sums)
    exp = knp.compare_image_stats(image_sums, ['Ref. band 1', 'Ref. band 2'])
    assert list(stats.keys()) == list(exp.keys()) == ['Ref. band 1', 'Ref. band 2', 'Mean']
    for band in exp:
        assert stats[band]['n'] == exp[band]['n'] == 500 and isinstance(stats[band]['n'], int)
        for key in ('r2', 'rmse', 'rrmse'):
            assert stats[band][key] == pytest.approx(exp[band][key], rel=1e-14)
    table = RasterCompare.stats_table(stats)
    assert 'RMSE' in table and 'Mean' in table and 'Ref. band 2' in table
    assert 'rRMSE' in RasterCompare.schema_table()

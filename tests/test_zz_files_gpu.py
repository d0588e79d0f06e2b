"""
GPU tier, file front-end: the reference's documented command-line flow (docs/cli.rst:44-72) through
GeoTIFF files -- `RasterFuse(src.tif, ref.tif).process(corr.tif, gain-blk-offset, (5, 5))`, then
`RasterCompare(corr.tif, landsat.tif).process()` -- against the numbers the reference publishes for it.

The files are written from tests/golden/docs_cli_ngi1.npz (the reference's test images do not travel to the GPU box).
"""
import json
import pathlib
import warnings

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from homonim_b200 import Affine, Model, RasterCompare, RasterFuse   # noqa: E402
from homonim_b200.files import create_out_postfix, create_param_filename   # noqa: E402
from homonim_b200.geotiff import GeoTiffReader, write_geotiff   # noqa: E402

GOLDEN = pathlib.Path(__file__).resolve().parent / 'golden'
GEOKEYS = ((1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, 32735), (), '')


def test_documented_file_flow_reproduces_published_table(tmp_path):
    meta = json.loads((GOLDEN / 'docs_cli_ngi1.json').read_text())
    with np.load(GOLDEN / 'docs_cli_ngi1.npz') as data:
        src, s2, l8 = data['src'], data['s2'], data['l8']
    src_path = write_geotiff(tmp_path / 'ngi_rgb_byte_1.tif', src, Affine(*meta['src_transform'][:6]), nodata=0,
                             geokeys=GEOKEYS, photometric='minisblack', blocksize=256)
    s2_path = write_geotiff(tmp_path / 'sentinel2_b432_byte.tif', s2, Affine(*meta['s2_transform'][:6]),
                            geokeys=GEOKEYS, descriptions=['B4', 'B3', 'B2'], photometric='rgb',
                            band_tags=[dict(center_wavelength=w) for w in (0.6645, 0.56, 0.4966)])
    # Landsat-8: the three matched bands among others, so that the wavelength matching has something to do
    l8_all = np.concatenate([l8[2:3] // 2, l8[2:3], l8[1:2], l8[0:1], l8[0:1] // 3])
    l8_path = write_geotiff(tmp_path / 'landsat8_byte.tif', l8_all, Affine(*meta['l8_transform'][:6]), nodata=0,
                            geokeys=GEOKEYS, descriptions=['SR_B1', 'SR_B2', 'SR_B3', 'SR_B4', 'SR_B5'],
                            band_tags=[dict(center_wavelength=w) for w in (0.443, 0.482, 0.562, 0.655, 0.865)])
    corr_path = tmp_path / ('ngi_rgb_byte_1' + create_out_postfix('ref', Model.gain_blk_offset, (5, 5)))
    param_path = create_param_filename(corr_path)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        with RasterFuse(src_path, s2_path) as fuse:
            assert fuse.src_bands == (1, 2, 3) and fuse.ref_bands == (1, 2, 3)
            fuse.process(corr_path, Model.gain_blk_offset, (5, 5), param_filename=param_path)
        rows = {}
        for key, path in (('source', src_path), ('corrected', corr_path)):
            with RasterCompare(path, l8_path) as compare:
                assert compare.ref_bands == (4, 3, 2)
                stats = compare.process()
                assert list(stats.keys()) == ['SR_B4', 'SR_B3', 'SR_B2', 'Mean']
                rows[key] = stats['Mean']
    for key in ('source', 'corrected'):
        r2, rmse, rrmse, n = meta['published'][key]
        got = rows[key]
        assert got['n'] == n, (key, got)
        for name, exp in (('r2', r2), ('rmse', rmse), ('rrmse', rrmse)):
            assert abs(got[name] - exp) <= 0.5e-3 + 1e-4 * abs(exp), (key, name, got[name], exp)
    with GeoTiffReader(corr_path) as im:
        assert im.dtype.name == 'float32' and im.count == 3 and im.descriptions == ['B4', 'B3', 'B2']
        assert im.tags()['FUSE_MODEL'] == 'gain_blk_offset' and im.tags()['FUSE_REF_FILE'] == 'sentinel2_b432_byte.tif'
    with GeoTiffReader(param_path) as im:
        assert im.count == 9 and im.descriptions[0] == 'B4_GAIN' and im.descriptions[8] == 'B2_R2'

"""
oracle/make_golden_docs.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Known answers PUBLISHED by the reference for its own test images: the "Summary over bands" table of
/root/reference/docs/cli.rst:58-72 -- `homonim fuse -m gain-blk-offset -k 5 5` of tests/data/source/ngi_rgb_byte_*.tif
with tests/data/reference/sentinel2_b432_byte.tif, then `homonim compare` of the source and corrected images with
tests/data/reference/landsat8_byte.tif.  Those numbers came out of the real pipeline (rasterio + GDAL), so they pin
the whole chain -- GDAL average down-sampling, the gain-blk-offset fit, GDAL cubic-spline up-sampling, the correction
and the RasterCompare sums -- including the GDAL steps that this container cannot run.

This script packs what is needed to re-run the ngi_rgb_byte_1.tif rows anywhere into tests/golden/docs_cli_ngi1.npz
(+ .json): the source image, and the reference bands matched to it (matched_pair.py:95-300: an RGB image without
wavelength metadata gets the standard red / green / blue wavelengths, which pair with Sentinel-2 B4 / B3 / B2 and
Landsat-8 SR_B4 / SR_B3 / SR_B2), cropped to the source's neighbourhood.  The images are read with oracle/tiff_min.py.

    python -m oracle.make_golden_docs
"""
import json
import pathlib
import sys

import numpy as np

REPO = pathlib.Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

from oracle import kernel_model_np as knp  # noqa: E402
from oracle.tiff_min import read_geotiff  # noqa: E402

DATA = pathlib.Path('/root/reference/tests/data')
GOLDEN_DIR = REPO / 'tests' / 'golden'
NAN = float('nan')
BAND_NAMES = ['SR_B4', 'SR_B3', 'SR_B2']

# docs/cli.rst:63-72: mean over bands of r2, RMSE, rRMSE, N per image
PUBLISHED = {
    'ngi_rgb_byte_1.tif': dict(source=(0.390, 93.517, 2.454, 28383), corrected=(0.924, 16.603, 0.489, 28383)),
    'ngi_rgb_byte_2.tif': dict(source=(0.488, 94.049, 2.380, 28166), corrected=(0.906, 15.590, 0.445, 28166)),
    'ngi_rgb_byte_3.tif': dict(source=(0.386, 88.610, 2.323, 27676), corrected=(0.881, 15.531, 0.456, 27676)),
    'ngi_rgb_byte_4.tif': dict(source=(0.607, 89.409, 2.412, 27342), corrected=(0.897, 15.702, 0.474, 27342)),
}


def matched_references():
    """ (Sentinel-2 bands, transform), (Landsat-8 bands, transform) in the order matched to the source's R, G, B. """
    s2 = read_geotiff(DATA / 'reference' / 'sentinel2_b432_byte.tif')
    l8 = read_geotiff(DATA / 'reference' / 'landsat8_byte.tif')
    assert s2['descriptions'] == ['B4', 'B3', 'B2'] and s2['nodata'] is None
    l8_idx = [l8['descriptions'].index(name) for name in BAND_NAMES]
    assert l8['nodata'] == 0.0
    return (s2['array'], s2['transform']), (np.ascontiguousarray(l8['array'][l8_idx]), l8['transform'])


def crop_to(array, transform, src_shape, src_transform, margin=3):
    """ `array` cropped to the whole pixels covering the source extent plus `margin` pixels (clipped to the raster). """
    hs, ws = src_shape
    res = transform[0]
    c0 = (src_transform[2] - transform[2]) / res
    c1 = (src_transform[2] + src_transform[0] * ws - transform[2]) / res
    r0 = (transform[5] - src_transform[5]) / res
    r1 = (transform[5] - (src_transform[5] + src_transform[4] * hs)) / res
    c0, r0 = max(int(np.floor(c0)) - margin, 0), max(int(np.floor(r0)) - margin, 0)
    c1, r1 = min(int(np.ceil(c1)) + margin, array.shape[2]), min(int(np.ceil(r1)) + margin, array.shape[1])
    out_tf = (transform[0], 0.0, transform[2] + c0 * res, 0.0, transform[4], transform[5] - r0 * res)
    return np.ascontiguousarray(array[:, r0:r1, c0:c1]), out_tf


def docs_rows(src, src_tf, s2, s2_tf, l8, l8_tf):
    """ The (source, corrected) rows of the table for one image, computed with the oracle. """
    rows = {}
    src_sums, corr_sums = [], []
    for b in range(3):
        src_sums.append(knp.compare_band(src[b], src_tf, 0.0, l8[b], l8_tf, 0.0, 'ref'))
        _, _, corr = knp.fuse_band_blocks(src[b], src_tf, 0.0, s2[b], s2_tf, NAN, 'gain-blk-offset', (5, 5))
        corr_sums.append(knp.compare_band(corr, src_tf, NAN, l8[b], l8_tf, 0.0, 'ref'))
    for key, sums in (('source', src_sums), ('corrected', corr_sums)):
        mean = knp.compare_image_stats(sums, BAND_NAMES)['Mean']
        rows[key] = (float(mean['r2']), float(mean['rmse']), float(mean['rrmse']), int(mean['n']))
    return rows


def check_rows(rows, published, name):
    for key in ('source', 'corrected'):
        got, exp = rows[key], published[key]
        assert got[3] == exp[3], (name, key, got, exp)
        assert all(round(g, 3) == e for g, e in zip(got[:3], exp[:3])), (name, key, got, exp)


def main():
    (s2, s2_tf), (l8, l8_tf) = matched_references()
    for name, published in PUBLISHED.items():
        src = read_geotiff(DATA / 'source' / name)
        assert src['nodata'] == 0.0
        rows = docs_rows(src['array'], src['transform'], s2, s2_tf, l8, l8_tf)
        check_rows(rows, published, name)
        print(name, rows)
    name = 'ngi_rgb_byte_1.tif'
    src = read_geotiff(DATA / 'source' / name)
    s2_c, s2_c_tf = crop_to(s2, s2_tf, src['array'].shape[1:], src['transform'])
    l8_c, l8_c_tf = crop_to(l8, l8_tf, src['array'].shape[1:], src['transform'])
    rows = docs_rows(src['array'], src['transform'], s2_c, s2_c_tf, l8_c, l8_c_tf)      # the crops change nothing
    check_rows(rows, PUBLISHED[name], name + ' (cropped references)')
    np.savez_compressed(GOLDEN_DIR / 'docs_cli_ngi1.npz', src=src['array'], s2=s2_c, l8=l8_c)
    meta = dict(citation='docs/cli.rst:58-72 (Summary over bands, ngi_rgb_byte_1.tif rows)', src_file=name,
                src_transform=list(src['transform']), src_nodata=0.0, s2_transform=list(s2_c_tf), s2_nodata=None,
                l8_transform=list(l8_c_tf), l8_nodata=0.0, band_names=BAND_NAMES, model='gain-blk-offset',
                kernel_shape=[5, 5], published=PUBLISHED[name], oracle=rows)
    (GOLDEN_DIR / 'docs_cli_ngi1.json').write_text(json.dumps(meta, indent=1, sort_keys=True))
    print('wrote', GOLDEN_DIR / 'docs_cli_ngi1.npz', src['array'].shape, s2_c.shape, l8_c.shape)


if __name__ == '__main__':
    main()

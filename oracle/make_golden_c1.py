"""
oracle/make_golden_c1.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Adds the REAL-DATA fixture of BASELINE.json's configs[0] to tests/golden/: a crop of band 1 of the reference's own test
images -- tests/data/source/ngi_rgb_byte_1.tif (5 m NGI aerial, uint8, nodata 0) against
tests/data/reference/sentinel2_b432_byte.tif (10 m Sentinel-2, uint8) -- corrected with gain-blk-offset 5x5,
proc_crs=ref, by the UNMODIFIED reference (imported through oracle/rasterio_stub; GDAL-backed steps served by
oracle/gdal_restate.py).  The two grids are mis-aligned by a fraction of a source pixel, which the synthetic fixtures
do not cover.  Runs only in the build container (needs /root/reference); the .npz it writes is committed.

    python -m oracle.make_golden_c1
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
DATA = pathlib.Path('/root/reference/tests/data')
NAN = float('nan')


def read_tiff(path, band):
    """ (array, (res, x0, y0)) of one band of a north-up GeoTIFF, via PIL (ModelPixelScale / ModelTiepoint tags). """
    from PIL import Image
    im = Image.open(path)
    scale, tie = im.tag_v2[33550], im.tag_v2[33922]
    a = np.array(im)
    if a.ndim == 3:
        a = a[:, :, band]
    elif band != 0:
        raise ValueError('PIL exposes only the first plane of a planar-configuration TIFF')
    return np.ascontiguousarray(a), (float(scale[0]), float(tie[3]), float(tie[4]))


def main():
    km, ra_mod, enums, rio = import_reference()
    Affine, CRS = rio.Affine, rio.crs.CRS
    RasterArray = ra_mod.RasterArray
    crs = CRS({'init': 'epsg:3857'})          # (placeholder: both files share one CRS; only the grids matter here)
    warnings.simplefilter('ignore')

    src_full, (sres, sx0, sy0) = read_tiff(DATA / 'source' / 'ngi_rgb_byte_1.tif', 0)
    ref_full, (rres, rx0, ry0) = read_tiff(DATA / 'reference' / 'sentinel2_b432_byte.tif', 0)
    src_full_tf = Affine(sres, 0, sx0, 0, -sres, sy0)
    ref_full_tf = Affine(rres, 0, rx0, 0, -rres, ry0)
    # a 512 x 512 crop of the source that includes part of its nodata border
    r0, c0, n = 0, 0, 512
    src = src_full[r0:r0 + n, c0:c0 + n].copy()
    src_tf = src_full_tf * Affine.translation(c0, r0)
    src_nodata = 0.0
    # the reference cropped to the source crop's neighbourhood (20 reference pixels of margin)
    inv = ~ref_full_tf
    ca, ra = inv * (src_tf * (0, 0))
    cb, rb = inv * (src_tf * (n, n))
    m = 20
    rc0, rr0 = max(int(np.floor(ca)) - m, 0), max(int(np.floor(ra)) - m, 0)
    rc1, rr1 = min(int(np.ceil(cb)) + m, ref_full.shape[1]), min(int(np.ceil(rb)) + m, ref_full.shape[0])
    ref = ref_full[rr0:rr1, rc0:rc1].astype('float32')
    ref_tf = ref_full_tf * Affine.translation(rc0, rr0)

    # reference window covering the source, source window covering that (raster_pair.py:292-296), boundless read of
    # the source (raster_array.py:175-199): explicit cropping / nodata padding, as in oracle/make_golden.py
    hs, ws = src.shape
    inv = ~ref_tf
    a0, b0 = inv * (src_tf * (0, 0))
    a1, b1 = inv * (src_tf * (ws, hs))
    wc0, wr0, wc1, wr1 = int(np.floor(a0)), int(np.floor(b0)), int(np.ceil(a1)), int(np.ceil(b1))
    ref_blk = ref[wr0:wr1, wc0:wc1].copy()
    ref_blk_tf = ref_tf * Affine.translation(wc0, wr0)
    sinv = ~src_tf
    sc0, sr0 = sinv * (ref_blk_tf * (0, 0))
    sc1, sr1 = sinv * (ref_blk_tf * (ref_blk.shape[1], ref_blk.shape[0]))
    pc0, pr0 = int(np.floor(sc0 + 1e-9)), int(np.floor(sr0 + 1e-9))
    pc1, pr1 = int(np.ceil(sc1 - 1e-9)), int(np.ceil(sr1 - 1e-9))
    src_blk = np.full((pr1 - pr0, pc1 - pc0), src_nodata, dtype='float32')
    src_blk[-pr0:-pr0 + hs, -pc0:-pc0 + ws] = src.astype('float32')
    src_blk_tf = src_tf * Affine.translation(pc0, pr0)

    model, kshape = 'gain-blk-offset', (5, 5)
    kmodel = km.RefSpaceModel(model, kshape, find_r2=True)
    mk = lambda: RasterArray(src_blk.copy(), crs, src_blk_tf, nodata=src_nodata)   # noqa: E731
    param_ra = kmodel.fit(mk(), RasterArray(ref_blk.copy(), crs, ref_blk_tf, nodata=NAN))
    corr_blk = kmodel.apply(mk(), param_ra).array
    corr = np.ascontiguousarray(corr_blk[-pr0:-pr0 + hs, -pc0:-pc0 + ws])

    name = 'refspace_c1_ngi_s2_gain-blk-offset_k5x5_uint8'
    np.savez_compressed(GOLDEN_DIR / f'{name}.npz', src=src, ref=ref, params=param_ra.array, corr=corr)
    index = json.loads((GOLDEN_DIR / 'index.json').read_text())
    index[name] = dict(kind='refspace', model=model, kernel_shape=list(kshape), find_r2=True, r2_inpaint_thresh=None,
                       mask_partial=False, src_nodata=src_nodata, ref_nodata='nan', src_transform=list(src_tf),
                       ref_transform=list(ref_tf), param_transform=list(ref_blk_tf),
                       source='real data: /root/reference/tests/data ngi_rgb_byte_1.tif band 1 x '
                              'sentinel2_b432_byte.tif band 1 (BASELINE.json configs[0]), 512 x 512 crop')
    (GOLDEN_DIR / 'index.json').write_text(json.dumps(index, indent=1, sort_keys=True))
    valid = np.isfinite(corr)
    print(f'{name}: src{src.shape} ref{ref.shape} params{param_ra.array.shape} corr{corr.shape}; '
          f'{valid.mean():.3f} of the crop valid; src grid offset vs ref grid: '
          f'{((sx0 - rx0) / sres) % 2:.3f}, {((ry0 - sy0) / sres) % 2:.3f} source pixels')


if __name__ == '__main__':
    main()

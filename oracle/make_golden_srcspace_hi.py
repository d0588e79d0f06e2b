"""
oracle/make_golden_srcspace_hi.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Additional SrcSpaceModel fixtures at larger reference-to-source ratios (the fixtures of oracle/make_golden.py use
ratio 2): ratio 4 and ratio 6 with a sub-pixel offset between the grids, produced by the UNMODIFIED reference (imported
through oracle/rasterio_stub; GDAL-backed steps served by oracle/gdal_restate.py).  They exercise the double-precision
reference up-sampling that feeds the source-resolution fit.

    python -m oracle.make_golden_srcspace_hi
"""
import json
import pathlib
import sys
import warnings

import numpy as np

REPO = pathlib.Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

from oracle.make_golden import upsample_pattern  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402

GOLDEN_DIR = REPO / 'tests' / 'golden'
NAN = float('nan')


def main():
    km, ra_mod, enums, rio = import_reference()
    Affine, CRS = rio.Affine, rio.crs.CRS
    RasterArray = ra_mod.RasterArray
    crs = CRS({'init': 'epsg:3857'})
    warnings.simplefilter('ignore')
    index = json.loads((GOLDEN_DIR / 'index.json').read_text())
    ref_tf = Affine(10, 0, 2000, 0, -10, 9000)
    cases = [('gain-offset', (5, 5), True, None, 4, (0.0, 0.0)), ('gain-blk-offset', (5, 5), False, None, 6, (1.5, 2.25))]
    for ci, (model, kshape, find_r2, thresh, ratio, shift) in enumerate(cases):
        rng = np.random.default_rng(400 + ci)
        hp, wp = 30, 26
        src, ref = upsample_pattern(rng, hp, wp, ratio, 0.3, 'float32', NAN)
        ref[3:5, 20:23] = NAN                                           # a hole in the reference
        src_tf = ref_tf * Affine.scale(1.0 / ratio) * Affine.translation(*shift)
        kmodel = km.SrcSpaceModel(model, kshape, find_r2=find_r2, r2_inpaint_thresh=thresh)
        param_ra = kmodel.fit(RasterArray(src.copy(), crs, src_tf, nodata=NAN),
                              RasterArray(ref.copy(), crs, ref_tf, nodata=NAN))
        corr = kmodel.apply(RasterArray(src.copy(), crs, src_tf, nodata=NAN), param_ra).array
        name = f'srcspace_hi{ci}_{model}_k{kshape[0]}x{kshape[1]}_r{ratio}'
        np.savez_compressed(GOLDEN_DIR / f'{name}.npz', src=src, ref=ref, params=param_ra.array, corr=corr)
        index[name] = dict(kind='srcspace', model=model, kernel_shape=list(kshape), find_r2=find_r2,
                           r2_inpaint_thresh=thresh, mask_partial=False, src_nodata='nan', ref_nodata='nan',
                           src_transform=list(src_tf), ref_transform=list(ref_tf))
        print(name, src.shape, ref.shape, param_ra.array.shape)
    (GOLDEN_DIR / 'index.json').write_text(json.dumps(index, indent=1, sort_keys=True))


if __name__ == '__main__':
    main()

"""
oracle/make_golden_convert.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Golden vectors for the output dtype conversion: RasterArray._convert_array_dtype
(/root/reference/homonim/raster_array.py:353-387) of the UNMODIFIED reference (imported through oracle/rasterio_stub)
on a float32 corrected plane with NaN nodata, exact .5 ties and out-of-range values.  Writes
tests/golden/convert_dtype.npz (not listed in index.json: it is not a fit / apply fixture).

    python -m oracle.make_golden_convert
"""
import pathlib
import sys
import warnings

import numpy as np

REPO = pathlib.Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

from oracle.ref_import import import_reference  # noqa: E402

CASES = [('uint8', 0), ('uint16', 0), ('uint16', 65535), ('int16', -32768), ('float32', -9999.0)]


def main():
    km, ra_mod, enums, rio = import_reference()
    warnings.simplefilter('ignore')
    rng = np.random.default_rng(8)
    n = 20000
    corr = rng.normal(300.0, 400.0, n).astype('float32')
    corr[::7] = np.round(corr[::7]) + 0.5                           # exact ties
    special = [-1e9, 1e9, 65535.4, 65535.5, 65536.0, -0.5, 0.5, 1.5, 2.5, 254.5, 255.5, 32767.5, -32768.5, 32766.5,
               -32767.5, 0.0, -0.0, 255.0, 65535.0]
    corr[5:5 + len(special)] = special
    corr[rng.integers(0, n, 1500)] = np.nan
    corr = corr.reshape(100, 200)
    out = dict(corr=corr)
    crs = rio.crs.CRS({'init': 'epsg:3857'})
    tf = rio.Affine(1, 0, 0, 0, -1, 0)
    for dtype, nodata in CASES:
        ra = ra_mod.RasterArray(corr.copy(), crs, tf, nodata=float('nan'))
        with np.errstate(all='ignore'):
            out[f'{dtype}_{nodata}'] = ra._convert_array_dtype(dtype, nodata=nodata)
    np.savez_compressed(REPO / 'tests' / 'golden' / 'convert_dtype.npz', **out)
    print({k: (v.dtype, v.shape) for k, v in out.items()})


if __name__ == '__main__':
    main()

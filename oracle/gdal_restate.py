"""
oracle/gdal_restate.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end for ``oracle/gdal_restate.c``: the CPU restatement of the GDAL calls the reference's hot path makes
through rasterio (``rasterio.warp.reproject`` -- /root/reference/homonim/raster_array.py:573-577 -- and
``rasterio.fill.fillnodata`` -- /root/reference/homonim/kernel_model.py:366).  GDAL / rasterio are not installed in
this image; see the C file's header for what is restated and how far each piece is pinned (average and cubic spline:
the reference's published known answers; fillnodata and nearest: PARITY UNPINNED).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
"""
import ctypes
import os
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_SRC = _HERE / 'gdal_restate.c'
_SO = _HERE / '_build' / 'libgdal_restate.so'
_lib = None

RESAMPLING_AVERAGE = 'average'
RESAMPLING_CUBIC_SPLINE = 'cubic_spline'
RESAMPLING_NEAREST = 'nearest'


def build(force: bool = False) -> pathlib.Path:
    """ Compile the C restatement with gcc (OpenMP). """
    if force or (not _SO.exists()) or (_SRC.exists() and _SO.stat().st_mtime < _SRC.stat().st_mtime):
        _SO.parent.mkdir(exist_ok=True)
        cmd = ['gcc', '-O2', '-fopenmp', '-fno-fast-math', '-ffp-contract=off', '-shared', '-fPIC', '-o', str(_SO),
               str(_SRC), '-lm']
        subprocess.run(cmd, check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(str(_SO))
        c_long, c_int, c_dbl, c_vp = ctypes.c_long, ctypes.c_int, ctypes.c_double, ctypes.c_void_p
        for name in ('gr_average_f32', 'gr_average_f64', 'gr_average_u8', 'gr_average_u16', 'gr_nearest'):
            getattr(lib, name).argtypes = [c_vp, c_long, c_long, c_int, c_dbl, c_vp, c_long, c_long, c_dbl, c_dbl,
                                           c_dbl, c_dbl]
            getattr(lib, name).restype = None
        lib.gr_cubic_spline_up.argtypes = [c_vp, c_long, c_long, c_long, c_int, c_dbl, c_vp, c_long, c_long, c_dbl,
                                           c_dbl, c_dbl, c_dbl]
        lib.gr_cubic_spline_up.restype = None
        lib.gr_fillnodata.argtypes = [c_vp, c_vp, c_long, c_long, c_dbl]
        lib.gr_fillnodata.restype = None
        _lib = lib
    return _lib


def grid_map(src_transform, dst_transform):
    """
    (sx, ox, sy, oy) mapping destination pixel-edge coordinates to source pixel-edge coordinates for two north-up,
    un-rotated geo-transforms given as 6-tuples / objects with attributes a..f (GDAL/affine order: x = a*col + c,
    y = e*row + f).
    """
    sa, sb, sc, sd, se, sf = [float(v) for v in tuple(src_transform)[:6]]
    da, db, dc, dd, de, df = [float(v) for v in tuple(dst_transform)[:6]]
    if sb != 0 or sd != 0 or db != 0 or dd != 0:
        raise NotImplementedError('rotated transforms are outside the restated path')
    if (sa > 0) != (da > 0) or (se > 0) != (de > 0):
        raise NotImplementedError('source and destination grids must have the same orientation')
    return da / sa, (dc - sc) / sa, de / se, (df - sf) / se


def _nodata_args(nodata):
    if nodata is None:
        return 0, 0.0
    return 1, float(nodata)


def reproject_array(src, src_transform, src_nodata, dst_shape, dst_transform, dst_nodata, resampling,
                    out_dtype='float32'):
    """
    Restated ``rasterio.warp.reproject`` for arrays on axis-aligned same-CRS grids.  ``src`` is 2D or 3D (bands
    first).  The destination is initialised with ``dst_nodata`` (rasterio ``init_dest_nodata=True``; 0 when
    ``dst_nodata`` is None) and returned as float32.
    """
    lib = _load()
    src = np.asarray(src)
    squeeze = src.ndim == 2
    src3 = src[None] if squeeze else src
    nb, hs, ws = src3.shape
    hd, wd = int(dst_shape[0]), int(dst_shape[1])
    sx, ox, sy, oy = grid_map(src_transform, dst_transform)
    has_nd, nd = _nodata_args(src_nodata)
    fill = 0.0 if dst_nodata is None else dst_nodata
    dst = np.full((nb, hd, wd), fill, dtype='float32')
    resampling = getattr(resampling, 'name', resampling)

    if resampling == RESAMPLING_AVERAGE:
        fn_by_dtype = {'float32': lib.gr_average_f32, 'float64': lib.gr_average_f64, 'uint8': lib.gr_average_u8,
                       'uint16': lib.gr_average_u16}
        if src3.dtype.name not in fn_by_dtype:
            src3 = src3.astype('float64')
        fn = fn_by_dtype[src3.dtype.name]
        for b in range(nb):
            plane = np.ascontiguousarray(src3[b])
            fn(plane.ctypes.data, hs, ws, has_nd, nd, dst[b].ctypes.data, hd, wd, sx, ox, sy, oy)
    elif resampling == RESAMPLING_CUBIC_SPLINE:
        if sx > 1.0 + 1e-12 or sy > 1.0 + 1e-12:
            raise NotImplementedError('cubic_spline is only restated for up-sampling')
        s64 = np.ascontiguousarray(src3, dtype='float64')
        lib.gr_cubic_spline_up(s64.ctypes.data, nb, hs, ws, has_nd, nd, dst.ctypes.data, hd, wd, sx, ox, sy, oy)
    elif resampling == RESAMPLING_NEAREST:
        for b in range(nb):
            plane = np.ascontiguousarray(src3[b], dtype='float64')
            lib.gr_nearest(plane.ctypes.data, hs, ws, has_nd, nd, dst[b].ctypes.data, hd, wd, sx, ox, sy, oy)
    else:
        raise NotImplementedError(f'resampling {resampling!r} is outside the restated path')
    dst = dst.astype(out_dtype, copy=False)
    return dst[0] if squeeze else dst


def fillnodata(image, mask, max_search_distance=100.0, smoothing_iterations=0):
    """ Restated ``rasterio.fill.fillnodata`` (GDALFillNodata); returns the filled float32 image. """
    if smoothing_iterations != 0:
        raise NotImplementedError('smoothing is outside the restated path')
    lib = _load()
    img = np.array(image, dtype='float32', copy=True, order='C')
    m = np.ascontiguousarray(np.asarray(mask) != 0, dtype='uint8')
    if img.ndim != 2 or m.shape != img.shape:
        raise ValueError('image and mask must be 2D with the same shape')
    lib.gr_fillnodata(img.ctypes.data, m.ctypes.data, img.shape[0], img.shape[1], float(max_search_distance))
    return img

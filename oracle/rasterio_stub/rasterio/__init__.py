"""
oracle/rasterio_stub/rasterio -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A name stub that lets the UNMODIFIED reference package (/root/reference/homonim) be imported in this container,
where rasterio / GDAL are not installed.  It provides just the names the reference touches at import time and on
the KernelModel / RefSpaceModel / SrcSpaceModel path:

  * plain data types: Affine, CRS, Window, Resampling, MaskFlags, ColorInterp;
  * ``rasterio.warp.reproject`` and ``rasterio.fill.fillnodata`` routed to the CPU restatements in
    ``oracle/gdal_restate.py`` (GDAL itself is absent -- see that file for the parity caveat);
  * everything that needs real file I/O raises NotImplementedError.

With it on sys.path (see oracle/ref_import.py) ``homonim.kernel_model.KernelModel.fit/apply`` executes the
reference's real cv2 + numpy code.  It is used only to generate the golden vectors under tests/golden/
(oracle/make_golden.py) and to cross-check the oracle port in this container; it never travels to the product path.
"""
import collections
import enum
import sys
import types

import numpy as np

__version__ = '0.0-oracle-stub'


class Affine(collections.namedtuple('Affine', 'a b c d e f')):
    """ Minimal 2D affine transform: x = a*col + b*row + c, y = d*col + e*row + f. """
    __slots__ = ()

    def __new__(cls, a, b, c, d, e, f, *_):
        return super().__new__(cls, float(a), float(b), float(c), float(d), float(e), float(f))

    @classmethod
    def identity(cls):
        return cls(1, 0, 0, 0, 1, 0)

    @classmethod
    def translation(cls, xoff, yoff):
        return cls(1, 0, xoff, 0, 1, yoff)

    @classmethod
    def scale(cls, sx, sy=None):
        return cls(sx, 0, 0, 0, sx if sy is None else sy, 0)

    def __mul__(self, other):
        if isinstance(other, Affine):
            sa, sb, sc, sd, se, sf = self
            oa, ob, oc, od, oe, of = other
            return Affine(sa * oa + sb * od, sa * ob + sb * oe, sa * oc + sb * of + sc,
                          sd * oa + se * od, sd * ob + se * oe, sd * oc + se * of + sf)
        x, y = other
        return (self.a * x + self.b * y + self.c, self.d * x + self.e * y + self.f)

    def __invert__(self):
        det = self.a * self.e - self.b * self.d
        ia, ib, id_, ie = self.e / det, -self.b / det, -self.d / det, self.a / det
        return Affine(ia, ib, -(ia * self.c + ib * self.f), id_, ie, -(id_ * self.c + ie * self.f))


class CRS:
    def __init__(self, data=None, **kwargs):
        self.data = dict(data or {}, **kwargs) if not isinstance(data, str) else {'init': data.lower()}

    @classmethod
    def from_epsg(cls, code):
        return cls({'init': f'epsg:{code}'})

    @classmethod
    def from_string(cls, s):
        return cls(s)

    def __eq__(self, other):
        return isinstance(other, CRS) and self.data == other.data

    def __hash__(self):
        return hash(tuple(sorted(self.data.items())))

    def __bool__(self):
        return True

    def to_wkt(self):
        return str(self.data)


class Resampling(enum.IntEnum):
    nearest = 0
    bilinear = 1
    cubic = 2
    cubic_spline = 3
    lanczos = 4
    average = 5
    mode = 6


class MaskFlags(enum.IntEnum):
    all_valid = 1
    per_dataset = 2
    alpha = 4
    nodata = 8


class ColorInterp(enum.IntEnum):
    undefined = 0
    gray = 1
    red = 3
    green = 4
    blue = 5
    alpha = 6


class Window(collections.namedtuple('Window', 'col_off row_off width height')):
    __slots__ = ()

    def toslices(self):
        return (slice(int(self.row_off), int(self.row_off + self.height)),
                slice(int(self.col_off), int(self.col_off + self.width)))

    def toranges(self):
        return ((self.row_off, self.row_off + self.height), (self.col_off, self.col_off + self.width))


def _window_transform(window, transform):
    return transform * Affine.translation(window.col_off, window.row_off)


def _window_bounds(window, transform):
    x0, y0 = transform * (window.col_off, window.row_off)
    x1, y1 = transform * (window.col_off + window.width, window.row_off + window.height)
    return (min(x0, x1), min(y0, y1), max(x0, x1), max(y0, y1))


class TransformMethodsMixin:
    pass


class WindowMethodsMixin:
    def window_transform(self, window):
        return _window_transform(window, self.transform)

    def window_bounds(self, window):
        return _window_bounds(window, self.transform)


def _needs_gdal(*args, **kwargs):
    raise NotImplementedError('this rasterio/GDAL call is outside the oracle stub (file I/O is out of scope)')


def _reproject(source, destination=None, src_transform=None, src_crs=None, src_nodata=None, dst_transform=None,
               dst_crs=None, dst_nodata=None, resampling=Resampling.nearest, num_threads=1, init_dest_nodata=True,
               **kwargs):
    """ rasterio.warp.reproject stand-in for in-memory arrays (axis-aligned, same CRS). """
    from oracle import gdal_restate
    if (src_crs is not None) and (dst_crs is not None) and (src_crs != dst_crs):
        raise NotImplementedError('CRS changes are outside the restated path')
    if dst_transform is None:
        dst_transform = src_transform
    out = gdal_restate.reproject_array(
        source, src_transform, src_nodata, destination.shape[-2:], dst_transform, dst_nodata,
        Resampling(resampling).name, out_dtype=destination.dtype
    )
    destination[...] = out
    return destination, dst_transform


def _fillnodata(image, mask=None, max_search_distance=100.0, smoothing_iterations=0):
    from oracle import gdal_restate
    return gdal_restate.fillnodata(image, mask, max_search_distance, smoothing_iterations)


def _can_cast_dtype(values, dtype):
    values = np.asarray(values)
    if np.issubdtype(np.dtype(dtype), np.floating):
        return True
    with np.errstate(invalid='ignore'):
        return bool(np.all(np.isfinite(values)) and np.all(values.astype(dtype) == values))


def _mod(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


class DatasetReader:
    pass


class _DatasetWriter:
    pass


crs = _mod('rasterio.crs', CRS=CRS)
enums = _mod('rasterio.enums', Resampling=Resampling, MaskFlags=MaskFlags, ColorInterp=ColorInterp)
transform = _mod('rasterio.transform', TransformMethodsMixin=TransformMethodsMixin, Affine=Affine)
windows = _mod(
    'rasterio.windows', Window=Window, WindowMethodsMixin=WindowMethodsMixin, transform=_window_transform,
    bounds=_window_bounds, get_data_window=_needs_gdal, intersect=_needs_gdal, union=_needs_gdal
)
warp = _mod(
    'rasterio.warp', reproject=_reproject, Resampling=Resampling, SUPPORTED_RESAMPLING=list(Resampling),
    calculate_default_transform=_needs_gdal, transform_bounds=_needs_gdal
)
errors = _mod('rasterio.errors', NotGeoreferencedWarning=type('NotGeoreferencedWarning', (UserWarning,), {}))
fill = _mod('rasterio.fill', fillnodata=_fillnodata)
io = _mod('rasterio.io', DatasetWriter=_DatasetWriter, DatasetReader=DatasetReader)
vrt = _mod('rasterio.vrt', WarpedVRT=type('WarpedVRT', (), {}))
dtypes = _mod('rasterio.dtypes', can_cast_dtype=_can_cast_dtype)
drivers = _mod('rasterio.drivers', raster_driver_extensions=_needs_gdal)
uint8, uint16, int16, float32, float64 = 'uint8', 'uint16', 'int16', 'float32', 'float64'
open = _needs_gdal
Env = _needs_gdal

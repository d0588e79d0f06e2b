"""
Rasterio-free geo-referencing primitives for the B200 kernel-model path.

The reference leans on ``rasterio.Affine`` / ``rasterio.crs.CRS`` (raster_array.py:29-30).  The hot path only needs
(i) equality of grids (kernel_model.py:430, 459) and (ii) the scale + sub-pixel offset between two north-up grids
of the same CRS, which is what the CUDA resamplers take as plain doubles.  Objects from rasterio / affine are accepted
anywhere a transform is expected (anything exposing a..f or iterating to >= 6 coefficients).
"""
from typing import NamedTuple, Tuple


class Affine(NamedTuple):
    """ x = a*col + b*row + c ;  y = d*col + e*row + f   (GDAL / affine coefficient order). """
    a: float
    b: float
    c: float
    d: float
    e: float
    f: float

    @classmethod
    def identity(cls) -> 'Affine':
        return cls(1.0, 0.0, 0.0, 0.0, 1.0, 0.0)

    @classmethod
    def translation(cls, xoff: float, yoff: float) -> 'Affine':
        return cls(1.0, 0.0, float(xoff), 0.0, 1.0, float(yoff))

    @classmethod
    def scale(cls, sx: float, sy: float = None) -> 'Affine':
        return cls(float(sx), 0.0, 0.0, 0.0, float(sx if sy is None else sy), 0.0)

    @classmethod
    def coerce(cls, transform) -> 'Affine':
        """ Build from an Affine, a rasterio/affine Affine, or any sequence of >= 6 coefficients. """
        if isinstance(transform, cls):
            return transform
        if all(hasattr(transform, k) for k in 'abcdef'):
            return cls(*(float(getattr(transform, k)) for k in 'abcdef'))
        coeffs = tuple(transform)
        if len(coeffs) < 6:
            raise TypeError('`transform` must provide 6 affine coefficients')
        return cls(*(float(v) for v in coeffs[:6]))

    def __mul__(self, other):
        if isinstance(other, Affine):
            return Affine(
                self.a * other.a + self.b * other.d, self.a * other.b + self.b * other.e,
                self.a * other.c + self.b * other.f + self.c,
                self.d * other.a + self.e * other.d, self.d * other.b + self.e * other.e,
                self.d * other.c + self.e * other.f + self.f,
            )
        col, row = other
        return (self.a * col + self.b * row + self.c, self.d * col + self.e * row + self.f)

    def __invert__(self) -> 'Affine':
        det = self.a * self.e - self.b * self.d
        if det == 0:
            raise ValueError('transform is not invertible')
        ia, ib, id_, ie = self.e / det, -self.b / det, -self.d / det, self.a / det
        return Affine(ia, ib, -(ia * self.c + ib * self.f), id_, ie, -(id_ * self.c + ie * self.f))

    @property
    def is_rectilinear(self) -> bool:
        return self.b == 0.0 and self.d == 0.0


class CRS:
    """ Opaque CRS label: the B200 path never re-projects between CRSs, it only compares them. """

    def __init__(self, definition=None):
        if isinstance(definition, CRS):
            definition = definition.definition
        self.definition = definition

    @classmethod
    def from_epsg(cls, code: int) -> 'CRS':
        return cls(f'EPSG:{int(code)}')

    def __eq__(self, other):
        return isinstance(other, CRS) and self.definition == other.definition

    def __hash__(self):
        return hash(str(self.definition))

    def __repr__(self):
        return f'CRS({self.definition!r})'


class GridMap(NamedTuple):
    """
    Axis-aligned mapping from DESTINATION pixel-edge coordinates to SOURCE pixel-edge coordinates:
    ``src_col = sx * dst_col + ox`` and ``src_row = sy * dst_row + oy``.  This is the whole geometry the CUDA
    resamplers need; it is passed through the C-ABI as four doubles.
    """
    sx: float
    ox: float
    sy: float
    oy: float


def grid_map(src_transform, dst_transform) -> GridMap:
    """ GridMap between two un-rotated grids of the same CRS and orientation. """
    s, d = Affine.coerce(src_transform), Affine.coerce(dst_transform)
    if not (s.is_rectilinear and d.is_rectilinear):
        raise NotImplementedError('rotated geo-transforms are not supported by the B200 kernel-model path')
    if (s.a > 0) != (d.a > 0) or (s.e > 0) != (d.e > 0):
        raise NotImplementedError('source and destination grids must have the same orientation')
    return GridMap(d.a / s.a, (d.c - s.c) / s.a, d.e / s.e, (d.f - s.f) / s.e)


def window_bounds(transform, height: int, width: int) -> Tuple[float, float, float, float]:
    """ (left, bottom, right, top) of an array with this transform (reference raster_array.py:276-279). """
    t = Affine.coerce(transform)
    x0, y0 = t * (0, 0)
    x1, y1 = t * (width, height)
    return (min(x0, x1), min(y0, y1), max(x0, x1), max(y0, y1))

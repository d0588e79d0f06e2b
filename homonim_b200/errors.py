"""
Exceptions and warnings raised on the kernel-model path.  The class names are the ones the reference raises in the same
situations (/root/reference/homonim/errors.py), so that ``except`` clauses written against homonim keep working; the
reference's file-format and block-size errors belong to components outside this path and are not defined here.
"""


class HomonimError(Exception):
    """ Base class of every exception this package raises itself. """


class NativeLibraryError(HomonimError):
    """ The sm_100a CUDA library is missing or failed to load, no CUDA device is present, or a native call returned an
    error.  There is no CPU fallback, so this is always fatal for the call. """


class IoError(HomonimError):
    """ A raster pair (``RasterFuse`` / ``RasterCompare``) was used before ``open()`` / outside its ``with`` block
    (reference raster_pair.py:271-278). """


class ImageContentError(HomonimError):
    """ The rasters cannot be fused: the reference does not cover the source, or has fewer bands than the source
    (reference raster_pair.py:160-192). """


class ImageProfileError(HomonimError):
    """ A ``RasterArray`` was given an incomplete or inconsistent georeferencing profile
    (reference raster_array.py:83-98). """


class HomonimWarning(RuntimeWarning):
    """ Base class of the warnings below; filter on it to silence the package. """


class ConfigWarning(HomonimWarning):
    """ A configuration value is allowed but not recommended, e.g. ``proc_crs`` set against the pixel-size ordering
    of source and reference (reference raster_pair.py:194-225). """


class BandMatchWarning(HomonimWarning):
    """ Something about the band selection / matching of two files was assumed rather than known, e.g. a three-band
    image without wavelength metadata taken to be RGB (reference matched_pair.py:118-176). """

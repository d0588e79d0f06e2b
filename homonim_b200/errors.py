"""
Exception and warning classes of the reference API (mirrors /root/reference/homonim/errors.py), plus the errors the
B200 path adds for its native library.
"""


class HomonimError(Exception):
    """ Root exception class. """


class UnsupportedImageError(HomonimError):
    """ Raised when an image cannot be handled. """


class ImageContentError(HomonimError):
    """ Raised when an image has insufficient coverage or bands. """


class BlockSizeError(HomonimError):
    """ Raised when the image block size is invalid. """


class ImageProfileError(HomonimError):
    """ Raised when an image profile is invalid. """


class ImageFormatError(HomonimError):
    """ Raised when an image format is invalid. """


class IoError(HomonimError):
    """ Raised when accessing unopened file(s). """


class NativeLibraryError(HomonimError):
    """ Raised when the sm_100a CUDA library is missing, fails to load, or a CUDA call fails (no CPU fallback). """


class HomonimWarning(RuntimeWarning):
    """ Homonim runtime warning. """


class BandMatchWarning(HomonimWarning):
    """ Warn about band matching issues. """


class ImageFormatWarning(HomonimWarning):
    """ Warn about image format issues. """


class ConfigWarning(HomonimWarning):
    """ Warn about configuration issues. """

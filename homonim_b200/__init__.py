"""
homonim_b200 -- B200-native (sm_100a CUDA) drop-in for the hot path of leftfield-geospatial/homonim: the per-pixel
sliding-kernel surface-reflectance model fit and correction (``KernelModel`` / ``RefSpaceModel`` / ``SrcSpaceModel``
``.fit()`` / ``.apply()``) behind ``RasterFuse.process``.

Host code is Python; all pixel arithmetic runs in hand-written CUDA kernels (``homonim_b200/csrc``) reached through
the C-ABI in ``include/homonim_b200.h``.  There is no CPU fallback.
"""
import logging

from homonim_b200.enums import Model, ProcCrs, Resampling
from homonim_b200.errors import NativeLibraryError
from homonim_b200.geometry import Affine, CRS
from homonim_b200.raster_array import RasterArray
from homonim_b200.kernel_model import KernelModel, RefSpaceModel, SrcSpaceModel
from homonim_b200.fuse import RasterFuse
from homonim_b200.compare import RasterCompare
from homonim_b200.geotiff import GeoTiffReader, write_geotiff

__version__ = '0.1.0'
logging.getLogger(__name__).addHandler(logging.NullHandler())

__all__ = ['Model', 'ProcCrs', 'Resampling', 'Affine', 'CRS', 'RasterArray', 'KernelModel', 'RefSpaceModel',
           'SrcSpaceModel', 'RasterFuse', 'RasterCompare', 'GeoTiffReader', 'write_geotiff', 'NativeLibraryError']

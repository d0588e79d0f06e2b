"""
oracle/ref_import.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Imports the UNMODIFIED reference package from /root/reference through the rasterio name stub.  Works only in the
build container (the GPU box has no /root/reference): used by oracle/make_golden.py and by the container-only
cross-check tests, which skip when the reference is absent.
"""
import importlib
import pathlib
import sys

REFERENCE_ROOT = pathlib.Path('/root/reference')
_STUB_DIR = pathlib.Path(__file__).resolve().parent / 'rasterio_stub'
_REPO_ROOT = pathlib.Path(__file__).resolve().parent.parent


def reference_available() -> bool:
    return (REFERENCE_ROOT / 'homonim' / 'kernel_model.py').exists()


def import_reference():
    """ Return the reference's (kernel_model, raster_array, enums) modules and the stub rasterio module. """
    if not reference_available():
        raise ImportError('/root/reference is not present (it only exists in the build container)')
    for p in (str(_REPO_ROOT), str(_STUB_DIR), str(REFERENCE_ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
    rasterio = importlib.import_module('rasterio')
    if getattr(rasterio, '__version__', '') != '0.0-oracle-stub':
        raise ImportError('a real rasterio is importable; the stub is not needed')
    kernel_model = importlib.import_module('homonim.kernel_model')
    raster_array = importlib.import_module('homonim.raster_array')
    enums = importlib.import_module('homonim.enums')
    return kernel_model, raster_array, enums, rasterio

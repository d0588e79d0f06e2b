"""
``RasterCompare`` -- accuracy statistics of a source raster against a reference (reference homonim/compare.py), with the
per-pixel work on the GPU: the re-projection onto the processing grid uses the kernel-model path's resampling kernels
and the masked sums are one HBM-bound reduction (``hb_compare_sums``).  Like ``RasterFuse`` it takes in-memory
``RasterArray`` objects (numpy arrays or CUDA tensors); file I/O is outside the B200 path.  There is no CPU fallback.
"""
from typing import Dict, List, Optional, Tuple

import numpy as np

from homonim_b200 import _native
from homonim_b200.enums import ProcCrs, Resampling
from homonim_b200.errors import IoError
from homonim_b200.fuse import RasterFuse, _band, _validate_threads
from homonim_b200.kernel_model import (NAN, _as_f32_plane, _call, _nodata_args, _require_torch, _resample_plane,
                                        _stream, _to_device)
from homonim_b200.raster_array import RasterArray

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None

SUM_KEYS = ('src_sum', 'ref_sum', 'src2_sum', 'ref2_sum', 'src_ref_sum', 'res2_sum', 'mask_sum')
""" Order of the 7 sums ``hb_compare_sums`` returns: the keys of the reference's ``sums_dict`` (compare.py:249-253). """


def compare_sums_device(src_t, src_nodata, ref_t, ref_nodata):
    """
    The masked sums of reference compare.py:243-253 for two device planes on one grid, as a float64 device tensor of
    7 values in ``SUM_KEYS`` order (enqueued on the current stream; nothing is copied to the host).
    """
    lib = _native.lib()
    if tuple(src_t.shape) != tuple(ref_t.shape):
        raise ValueError(f'source {tuple(src_t.shape)} and reference {tuple(ref_t.shape)} planes differ in shape')
    src_t, ref_t = _as_f32_plane(src_t, src_nodata).contiguous(), _as_f32_plane(ref_t, ref_nodata).contiguous()
    s_has, s_nd = _nodata_args(src_nodata)
    r_has, r_nd = _nodata_args(ref_nodata)
    sums = torch.empty(len(SUM_KEYS), dtype=torch.float64, device=src_t.device)
    ws_bytes = lib.hb_compare_sums_workspace_bytes()
    work = torch.empty(ws_bytes, dtype=torch.uint8, device=src_t.device)
    _call('hb_compare_sums', src_t.data_ptr(), s_has, s_nd, ref_t.data_ptr(), r_has, r_nd, int(src_t.numel()),
          sums.data_ptr(), work.data_ptr(), ws_bytes, _stream())
    return sums


class RasterCompare:
    schema = dict(
        r2=dict(abbrev='r\N{SUPERSCRIPT TWO}', description="Pearson's correlation coefficient squared"),
        rmse=dict(abbrev='RMSE', description='Root Mean Square Error'),
        rrmse=dict(abbrev='rRMSE', description='Relative RMSE (RMSE/mean(ref))'),
        n=dict(abbrev='N', description='Number of pixels'),
    )
    """ The statistics returned by :meth:`RasterCompare.process` (reference compare.py:83-88). """

    def __init__(self, src: RasterArray, ref: RasterArray, proc_crs: ProcCrs = ProcCrs.auto,
                 src_bands: Optional[List[int]] = None, ref_bands: Optional[List[int]] = None, force: bool = False,
                 band_names: Optional[List[str]] = None):
        """
        Compare a source raster with a reference (reference compare.py:38-81).  Arguments as for
        :class:`~homonim_b200.fuse.RasterFuse` (GeoTIFF file names or RasterArrays); ``band_names`` overrides the band
        descriptions read from the files (compare.py:171-175; default for RasterArrays: ``'Ref. band N'``).
        """
        self._pair = RasterFuse(src, ref, proc_crs=proc_crs, src_bands=src_bands, ref_bands=ref_bands, force=force)
        if band_names is not None and len(band_names) < len(self._pair.src_bands):
            raise ValueError('`band_names` must name every compared band')
        self._band_names = list(band_names) if band_names is not None else None
        self._closed = True

    # ---- pair properties and context management (reference raster_pair.py:104-134, 271-311) -------------------------
    @property
    def proc_crs(self) -> ProcCrs:
        return self._pair.proc_crs

    @property
    def src_bands(self) -> Tuple[int, ...]:
        return self._pair.src_bands

    @property
    def ref_bands(self) -> Tuple[int, ...]:
        return self._pair.ref_bands

    @property
    def closed(self) -> bool:
        return self._closed

    def open(self):
        self._closed = False

    def close(self):
        self._closed = True

    def __enter__(self):
        self.open()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.close()

    def _assert_open(self):
        if self._closed:
            raise IoError('The raster pair has not been opened: use it in a `with` block, or call `open()`.')

    # ---- configuration and tables (reference compare.py:90-131, 189-211) --------------------------------------------
    @staticmethod
    def schema_table() -> str:
        """ Table describing the statistics (reference compare.py:90-94). """
        from tabulate import tabulate
        headers = {key: key.upper() for key in list(RasterCompare.schema.values())[0].keys()}
        return tabulate(RasterCompare.schema.values(), headers=headers)

    @staticmethod
    def create_config(threads: int = 0, max_block_mem: float = 512, downsampling: Resampling = Resampling.average,
                      upsampling: Resampling = Resampling.cubic_spline) -> Dict:
        """ Configuration dictionary with defaults (reference compare.py:96-131).  ``threads`` and ``max_block_mem``
        are validated and kept for API compatibility: a band is one block on the GPU. """
        return dict(threads=_validate_threads(threads), max_block_mem=max_block_mem,
                    downsampling=Resampling.coerce(downsampling), upsampling=Resampling.coerce(upsampling))

    def _get_resampling(self, from_res, to_res, **kwargs) -> Resampling:
        """ Reference compare.py:133-139. """
        config = self.create_config(**kwargs)
        down = np.prod(np.abs(from_res)) <= np.prod(np.abs(to_res))
        return config['downsampling'] if down else config['upsampling']

    @staticmethod
    def stats_table(stats_dict: Dict[str, Dict], key_header: str = 'band') -> str:
        """ Table string of the statistics returned by :meth:`process` (reference compare.py:189-211). """
        from tabulate import tabulate
        stats_list = [dict(**{key_header: key}, **val) for key, val in stats_dict.items()]
        headers = {k: RasterCompare.schema[k]['abbrev'] if k in RasterCompare.schema else str.capitalize(k)
                   for k in list(stats_list[0].keys())}
        return tabulate(stats_list, headers=headers, floatfmt='.3f', stralign='right')

    # ---- statistics -------------------------------------------------------------------------------------------------
    @staticmethod
    def _band_stats(src_sum: float = 0, ref_sum: float = 0, src2_sum: float = 0, ref2_sum: float = 0,
                    src_ref_sum: float = 0, res2_sum: float = 0, mask_sum: float = 0) -> Dict:
        """ Statistics of one band from its sums (reference compare.py:145-163, same expressions in float64). """
        with np.errstate(divide='ignore', invalid='ignore'):
            mask_sum = np.float64(mask_sum)
            src_mean, ref_mean = np.float64(src_sum) / mask_sum, np.float64(ref_sum) / mask_sum
            pcc_num = src_ref_sum - (mask_sum * src_mean * ref_mean)
            pcc_den = (np.sqrt(src2_sum - (mask_sum * (src_mean ** 2))) *
                       np.sqrt(ref2_sum - (mask_sum * (ref_mean ** 2))))
            pcc = pcc_num / pcc_den
            rmse = np.sqrt(res2_sum / mask_sum)
            rrmse = rmse / ref_mean
        return dict(r2=float(pcc ** 2), rmse=float(rmse), rrmse=float(rrmse), n=int(mask_sum))

    def _band_name(self, band_i: int) -> str:
        if self._band_names is not None:
            return self._band_names[band_i]
        files = self._pair._files
        if files is not None:            # the reference band's description, else the source band's (compare.py:171-173)
            name = files.ref_descriptions[band_i] or files.src_descriptions[band_i]
            if name:
                return name
        return f'Ref. band {self.ref_bands[band_i]}'                        # compare.py:174

    def _get_image_stats(self, image_sums: List[Dict]) -> Dict[str, Dict]:
        """ Per-band statistics and their mean over the bands (reference compare.py:142-187). """
        image_stats, sum_over_bands = {}, {}
        for band_i, band_sums in enumerate(image_sums):
            band_stats = self._band_stats(**band_sums)
            image_stats[self._band_name(band_i)] = band_stats
            sum_over_bands = {k: sum_over_bands.get(k, 0) + v for k, v in band_stats.items()}
        image_stats['Mean'] = {k: int(v / len(image_sums)) if isinstance(v, int) else (v / len(image_sums))
                               for k, v in sum_over_bands.items()}
        return image_stats

    def _band_sums_device(self, band_i: int, **kwargs):
        """ `get_block_sums` of reference compare.py:232-256 for one band (read as a single block): re-project onto
        the processing grid, then the masked sums -- returns the 7 sums as a float64 device tensor. """
        src_ra = _band(self._pair._src, self._pair._src_bands[band_i])
        ref_ra = self._pair._ref_block(band_i)
        src_t, ref_t = _to_device(src_ra.array), _to_device(ref_ra.array)
        src_nodata, ref_nodata = src_ra.nodata, ref_ra.nodata
        if self.proc_crs == ProcCrs.ref:
            resampling = self._get_resampling(src_ra.res, ref_ra.res, **kwargs)              # :237
            src_t = _resample_plane(src_t, src_ra.transform, src_nodata, ref_ra.shape, ref_ra.transform,
                                    resampling)                                              # :238
            src_nodata = NAN
        else:
            resampling = self._get_resampling(ref_ra.res, src_ra.res, **kwargs)              # :240
            ref_t = _resample_plane(ref_t, ref_ra.transform, ref_nodata, src_ra.shape, src_ra.transform,
                                    resampling)                                              # :241
            ref_nodata = NAN
        return compare_sums_device(src_t, src_nodata, ref_t, ref_nodata)

    def process(self, **kwargs) -> Dict[str, Dict]:
        """
        Compare source and reference (reference compare.py:213-277): per band, both rasters are brought onto the
        processing grid and reduced to 7 sums on the GPU; the statistics are formed from the sums on the host.
        Keyword arguments as for :meth:`create_config`.
        """
        self._assert_open()
        _require_torch()
        self.create_config(**kwargs)
        sums_dev = [self._band_sums_device(band_i, **kwargs) for band_i in range(len(self.src_bands))]
        sums = torch.stack(sums_dev).cpu().numpy()                  # one device -> host read for all bands
        image_sums = [dict(zip(SUM_KEYS, (float(v) for v in row))) for row in sums]
        return self._get_image_stats(image_sums)

"""
oracle/make_golden_compare.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Golden vectors for the RasterCompare accuracy statistics: `RasterCompare.process` of the UNMODIFIED
reference (/root/reference/homonim/compare.py:199-277, imported through oracle/rasterio_stub) run on in-memory rasters.
Only the file reader is replaced (a subclass overrides `read` / `block_pairs`, which wrap rasterio dataset I/O, and
returns what the reader returns: the boundless, nodata-padded source window on whole reference pixels); the
re-projection, masking, sums and statistics are the reference's own code (GDAL's warper served by
oracle/gdal_restate.py, like every other fixture here).  Writes tests/golden/compare_stats.npz + compare_stats.json.

    python -m oracle.make_golden_compare
"""
import importlib
import json
import pathlib
import sys
import types
import warnings

import numpy as np

REPO = pathlib.Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

from oracle.ref_import import import_reference  # noqa: E402

GOLDEN_DIR = REPO / 'tests' / 'golden'
NAN = float('nan')
BAND_NAMES = ['B4', 'B3', 'B2']


def make_pair(rng, hp, wp, ratio, shift):
    """ 3-band float32 source (fine grid) and reference (coarse grid) with holes; source ~ linear in the reference. """
    yy, xx = np.mgrid[0:hp, 0:wp].astype('float32')
    ref = np.stack([
        (900 + 350 * np.sin(xx / (3.0 + b)) * np.cos(yy / (4.0 + b)) + rng.normal(0, 60, (hp, wp))).astype('float32')
        for b in range(3)
    ])
    src = np.repeat(np.repeat(ref, ratio, axis=1), ratio, axis=2)
    src = ((0.6 + 0.1 * np.arange(3)[:, None, None]) * src + 40 + rng.normal(0, 45, src.shape)).astype('float32')
    # the source reaches into the last reference row and column, so that the reference window the reader would crop
    # to (raster_pair.py:292-296) is the whole reference raster
    hs, ws = hp * ratio - int(np.ceil(shift[1])) - 2, wp * ratio - int(np.ceil(shift[0])) - 1
    src = np.ascontiguousarray(src[:, :hs, :ws])
    src[:, 5:9, 11:30] = NAN
    src[1, 20:22, 3:6] = NAN
    ref[:, 7:9, 2:4] = NAN
    return src, ref


def main():
    km, ra_mod, enums, rio = import_reference()
    cmp_mod = importlib.import_module('homonim.compare')
    rp_mod = importlib.import_module('homonim.raster_pair')
    utils_mod = importlib.import_module('homonim.utils')
    Affine, CRS, Window = rio.Affine, rio.crs.CRS, rio.windows.Window
    RasterArray = ra_mod.RasterArray
    crs = CRS({'init': 'epsg:3857'})
    warnings.simplefilter('ignore')

    class MemoryCompare(cmp_mod.RasterCompare):
        """ The reference's RasterCompare reading from in-memory rasters instead of rasterio datasets. """

        def __init__(self, src, src_tf, ref, ref_tf, proc_crs):
            self._src, self._src_tf, self._ref, self._ref_tf = src, src_tf, ref, ref_tf
            self._proc_crs = proc_crs
            self._src_bands, self._ref_bands = (1, 2, 3), (1, 2, 3)
            self._src_im = types.SimpleNamespace(descriptions=[None] * 3, closed=False)
            self._ref_im = types.SimpleNamespace(descriptions=list(BAND_NAMES), closed=False)
            self.image_sums = None
            self.src_windows = {}

        def read(self, block_pair):
            """ What RasterPairReader.read returns for a whole-image block: the windows of RasterPairReader.open
            (raster_pair.py:292-296: the reference window covering the source, expanded to whole reference pixels;
            the source window covering THAT, expanded to whole source pixels), read boundlessly, i.e. filled with
            nodata beyond the raster (raster_array.py:175-199). """
            b = block_pair.band_i
            s_tf, r_tf = self._src_tf, self._ref_tf
            hs, ws = self._src.shape[-2:]
            left, top = s_tf.c, s_tf.f
            right, bottom = s_tf.c + s_tf.a * ws, s_tf.f + s_tf.e * hs
            ref_win = utils_mod.expand_window_to_grid(Window(
                (left - r_tf.c) / r_tf.a, (top - r_tf.f) / r_tf.e, (right - left) / r_tf.a, (bottom - top) / r_tf.e))
            assert (ref_win.col_off, ref_win.row_off, ref_win.width, ref_win.height) == (
                0, 0, self._ref.shape[2], self._ref.shape[1]), ref_win
            r_left, r_top = r_tf.c, r_tf.f
            r_right, r_bottom = r_tf.c + r_tf.a * ref_win.width, r_tf.f + r_tf.e * ref_win.height
            src_win = utils_mod.expand_window_to_grid(Window(
                (r_left - s_tf.c) / s_tf.a, (r_top - s_tf.f) / s_tf.e, (r_right - r_left) / s_tf.a,
                (r_bottom - r_top) / s_tf.e))
            array = np.full((int(src_win.height), int(src_win.width)), NAN, 'float32')
            r0, c0 = -int(src_win.row_off), -int(src_win.col_off)
            array[r0:r0 + hs, c0:c0 + ws] = self._src[b]
            win_tf = s_tf * Affine.translation(int(src_win.col_off), int(src_win.row_off))
            self.src_windows[b] = [int(src_win.col_off), int(src_win.row_off), int(src_win.width), int(src_win.height)]
            return (RasterArray(array, crs, win_tf, nodata=NAN),
                    RasterArray(self._ref[b].copy(), crs, r_tf, nodata=NAN))

        def block_pairs(self, overlap=(0, 0), max_block_mem=np.inf):
            sw = Window(0, 0, self._src.shape[2], self._src.shape[1])
            rw = Window(0, 0, self._ref.shape[2], self._ref.shape[1])
            for b in range(3):
                yield rp_mod.BlockPair(b, sw, rw, sw, rw, True)

        def _get_image_stats(self, image_sums):          # record the accumulated sums, then the reference's own code
            self.image_sums = [{k: float(v) for k, v in d.items()} for d in image_sums]
            return super()._get_image_stats(image_sums)

    ref_tf = Affine(10, 0, 2000, 0, -10, 9000)
    cases = [('ref', 4, (0.0, 0.0)), ('src', 3, (1.5, 2.25)), ('ref', 5, (2.0, 1.0))]
    arrays, meta = {}, {}
    for ci, (proc, ratio, shift) in enumerate(cases):
        rng = np.random.default_rng(900 + ci)
        src, ref = make_pair(rng, 40, 36, ratio, shift)
        src_tf = ref_tf * Affine.scale(1.0 / ratio) * Affine.translation(*shift)
        cmp = MemoryCompare(src, src_tf, ref, ref_tf, enums.ProcCrs(proc))
        stats = cmp.process(threads=1)
        arrays[f'src{ci}'], arrays[f'ref{ci}'] = src, ref
        meta[f'case{ci}'] = dict(proc_crs=proc, src_transform=list(src_tf), ref_transform=list(ref_tf),
                                 stats={k: {kk: float(vv) for kk, vv in v.items()} for k, v in stats.items()},
                                 image_sums=cmp.image_sums, src_window=cmp.src_windows[0])
        print(ci, proc, ratio, src.shape, ref.shape, {k: {kk: round(float(vv), 5) for kk, vv in v.items()}
                                                       for k, v in stats.items()})
    np.savez_compressed(GOLDEN_DIR / 'compare_stats.npz', **arrays)
    (GOLDEN_DIR / 'compare_stats.json').write_text(json.dumps(meta, indent=1, sort_keys=True))


if __name__ == '__main__':
    main()

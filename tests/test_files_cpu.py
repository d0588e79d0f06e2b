"""
CPU tier: the GeoTIFF reader / writer, band matching and the file front-end of RasterFuse / RasterCompare
(SURVEY.md 8f-4).  No GPU: pixels are only moved, never computed on.
"""
import pathlib
import warnings

import numpy as np
import pytest

from homonim_b200 import Affine, Model, ProcCrs, RasterArray, RasterCompare, RasterFuse
from homonim_b200.errors import BandMatchWarning
from homonim_b200.files import FilePair, create_out_postfix, create_param_filename
from homonim_b200.geotiff import GeoTiffReader, write_geotiff
from homonim_b200.matched_pair import band_info, match_bands

REF_DATA = pathlib.Path('/root/reference/tests/data')
needs_reference = pytest.mark.skipif(not (REF_DATA / 'source' / 'ngi_rgb_byte_1.tif').exists(),
                                     reason='container-only: reads the reference test images from /root/reference')
GEOKEYS = ((1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, 3857), (), '')
TF = Affine(5.0, 0.0, -57129.5, 0.0, -5.0, -3723906.75)


@pytest.mark.parametrize('dtype, count, interleave, compress, blocksize, bigtiff', [
    ('uint8', 3, 'band', 'deflate', 64, None),
    ('uint16', 4, 'pixel', 'deflate', 32, None),
    ('float32', 1, 'band', None, 16, None),
    ('int16', 2, 'band', 'deflate', 48, None),
    ('float32', 9, 'band', 'deflate', 16, True),
    ('float64', 2, 'pixel', None, 32, True),
])
def test_geotiff_round_trip(tmp_path, dtype, count, interleave, compress, blocksize, bigtiff):
    rng = np.random.default_rng(1)
    array = rng.normal(100, 40, (count, 70, 53)).astype(dtype)
    nodata = float('nan') if dtype.startswith('float') else 0
    path = write_geotiff(tmp_path / 'x.tif', array, TF, nodata=nodata, descriptions=[f'b{i}' for i in range(count)],
                         tags=dict(FUSE_MODEL='gain & <offset>'), geokeys=GEOKEYS, compress=compress,
                         band_tags=[dict(center_wavelength=0.5 + i) for i in range(count)], blocksize=blocksize,
                         interleave=interleave, bigtiff=bigtiff)
    with pytest.raises(FileExistsError):
        write_geotiff(path, array, TF)
    with GeoTiffReader(path) as im:
        assert (im.count, im.height, im.width, im.dtype.name) == (count, 70, 53, dtype)
        assert im.transform == TF and im.block_shape == (blocksize, blocksize) and im.tiled
        assert (np.isnan(im.nodata) if dtype.startswith('float') else im.nodata == 0)
        assert im.descriptions == [f'b{i}' for i in range(count)]
        assert im.tags() == dict(FUSE_MODEL='gain & <offset>') and im.tags(2 if count > 1 else 1) == \
            dict(center_wavelength=str(0.5 + (1 if count > 1 else 0)))
        assert im.geokeys[0] == GEOKEYS[0] and im.crs is not None
        assert np.array_equal(im.read(), array, equal_nan=True)
        assert np.array_equal(im.read(1), array[0], equal_nan=True)                 # scalar index -> 2D
        window = im.read([count, 1], window=(10, 20, 30, 25))                       # band order follows `indexes`
        assert np.array_equal(window, array[[count - 1, 0], 20:45, 10:40], equal_nan=True)
        # boundless window: nodata (or 0) beyond the raster (raster_array.py:175-199)
        out = im.read(1, window=(-4, -3, 20, 10))
        assert np.array_equal(out[3:, 4:], array[0, :7, :16], equal_nan=True)
        edge = out[:3, :]
        assert np.all(np.isnan(edge)) if dtype.startswith('float') else np.all(edge == 0)
        pinned = im.read([1], pinned=True)
        assert np.array_equal(np.asarray(pinned), array[:1], equal_nan=True)
        assert im._big == bool(bigtiff)


def test_geotiff_reads_strips_predictor_and_big_endian(tmp_path):
    """ Hand-assembled big-endian, stripped, chunky, deflate + horizontal-predictor TIFF (what this writer never
    produces but other writers do). """
    import struct
    import zlib
    rng = np.random.default_rng(2)
    array = rng.integers(0, 60000, (2, 11, 7)).astype('uint16')
    rows_per_strip = 4
    strips = []
    for y0 in range(0, 11, rows_per_strip):
        chunk = np.moveaxis(array[:, y0:y0 + rows_per_strip, :], 0, 2).astype('>u2')     # rows, cols, samples
        diff = chunk.copy()
        diff[:, 1:, :] = chunk[:, 1:, :] - chunk[:, :-1, :]                              # horizontal differencing
        strips.append(zlib.compress(diff.tobytes()))
    entries = []
    data = b''
    base = 8

    def put(tag, typ, values):
        fmt = {3: 'H', 4: 'I', 12: 'd'}[typ]
        entries.append((tag, typ, len(values), struct.pack('>' + fmt * len(values), *values)))

    offsets, pos = [], base
    for s in strips:
        offsets.append(pos)
        pos += len(s)
    put(256, 3, [7]); put(257, 3, [11]); put(258, 3, [16, 16]); put(259, 3, [8]); put(262, 3, [1])
    put(273, 4, offsets); put(277, 3, [2]); put(278, 3, [rows_per_strip]); put(279, 4, [len(s) for s in strips])
    put(284, 3, [1]); put(317, 3, [2]); put(338, 3, [0]); put(339, 3, [1, 1])
    put(33550, 12, [2.0, 2.0, 0.0]); put(33922, 12, [0.0, 0.0, 0.0, 100.0, 200.0, 0.0])
    body = b''.join(strips)
    extra_pos = base + len(body)
    extra = b''
    ifd = b''
    for tag, typ, cnt, packed in entries:
        if len(packed) <= 4:
            value = packed.ljust(4, b'\x00')
        else:
            value = struct.pack('>I', extra_pos + len(extra))
            extra += packed + (b'\x00' if len(packed) % 2 else b'')
        ifd += struct.pack('>HHI', tag, typ, cnt) + value
    ifd_pos = extra_pos + len(extra)
    blob = b'MM' + struct.pack('>HI', 42, ifd_pos) + body + extra + struct.pack('>H', len(entries)) + ifd + \
        struct.pack('>I', 0)
    path = tmp_path / 'be.tif'
    path.write_bytes(blob)
    with GeoTiffReader(path) as im:
        assert not im.tiled and im.block_shape == (rows_per_strip, 7) and im.planar == 1 and im.predictor == 2
        assert im.transform == Affine(2.0, 0.0, 100.0, 0.0, -2.0, 200.0) and im.nodata is None and im.crs is None
        assert np.array_equal(im.read(), array)
        assert np.array_equal(im.read(2, window=(2, 3, 4, 6)), array[1, 3:9, 2:6])


def test_geotiff_rejects_what_it_cannot_decode(tmp_path):
    path = write_geotiff(tmp_path / 'x.tif', np.zeros((1, 16, 16), 'uint8'), TF)
    blob = bytearray(path.read_bytes())
    with GeoTiffReader(path) as im:
        assert im.compression == 8
    # patch the Compression tag (259) to LZW (5)
    import struct
    ifd, = struct.unpack('<I', blob[4:8])
    n, = struct.unpack('<H', blob[ifd:ifd + 2])
    for i in range(n):
        pos = ifd + 2 + 12 * i
        if struct.unpack('<H', blob[pos:pos + 2])[0] == 259:
            blob[pos + 8:pos + 10] = struct.pack('<H', 5)
    bad = tmp_path / 'lzw.tif'
    bad.write_bytes(bytes(blob))
    with pytest.raises(NotImplementedError):
        GeoTiffReader(bad)
    with pytest.raises(NotImplementedError):
        write_geotiff(tmp_path / 'y.tif', np.zeros((1, 4, 4), 'uint8'), TF, compress='jpeg')
    with pytest.raises(NotImplementedError):
        write_geotiff(tmp_path / 'z.tif', np.zeros((1, 4, 4), 'uint8'), Affine(1, 0.1, 0, 0, -1, 0))


class _FakeImage:
    """ The attributes band matching looks at. """
    def __init__(self, name, wavelengths, colorinterp=None, descriptions=None):
        self.name, self.count = name, len(wavelengths)
        self._wl = wavelengths
        self.colorinterp = colorinterp or ['undefined'] * self.count
        self.descriptions = descriptions or [None] * self.count

    def tags(self, bidx):
        w = self._wl[bidx - 1]
        return {} if w is None else dict(center_wavelength=str(w))


def test_band_matching_rules():
    """ matched_pair.py:95-342: wavelength matching, RGB assumption, helper bands, file-order fallback, errors. """
    landsat = _FakeImage('l8.tif', [0.443, 0.482, 0.562, 0.655, 0.865, None, None],
                         descriptions=['SR_B1', 'SR_B2', 'SR_B3', 'SR_B4', 'SR_B5', 'FILL_MASK', 'CLOUD_DIST'])
    rgb = _FakeImage('rgb.tif', [None, None, None])
    with pytest.warns(BandMatchWarning, match='Assuming image is RGB'):
        assert match_bands(rgb, landsat) == ((1, 2, 3), (4, 3, 2))
    bgr = _FakeImage('bgr.tif', [None, None, None], colorinterp=['blue', 'green', 'red'])
    with pytest.warns(BandMatchWarning, match='Assigning standard'):
        assert match_bands(bgr, landsat) == ((1, 2, 3), (2, 3, 4))
    # mask / distance / alpha bands are never used
    assert band_info(landsat)[0] == [1, 2, 3, 4, 5]
    rgba = _FakeImage('rgba.tif', [None] * 4, colorinterp=['red', 'green', 'blue', 'alpha'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        assert band_info(rgba)[0] == [1, 2, 3]
        with pytest.raises(ValueError, match='alpha'):
            band_info(rgba, [4])
        with pytest.raises(ValueError, match='invalid'):
            band_info(rgba, [7])
    # no wavelengths anywhere, equal counts: file order
    a, b = _FakeImage('a.tif', [None] * 4), _FakeImage('b.tif', [None] * 4)
    assert match_bands(a, b) == ((1, 2, 3, 4), (1, 2, 3, 4))
    # user subsets keep their order; each reference band is used once
    modis = _FakeImage('modis.tif', [0.645, 0.8585, 0.469, 0.555, 1.24, 1.64, 2.13])
    assert match_bands(landsat, modis, src_bands=[4, 3, 2]) == ((4, 3, 2), (1, 4, 3))
    with pytest.raises(ValueError, match='fewer bands'):
        match_bands(modis, _FakeImage('two.tif', [0.5, 0.6]))
    # a pairing further apart than 10 % is an error
    with pytest.raises(ValueError, match='could not be auto-matched'):
        match_bands(_FakeImage('swir.tif', [2.2]), _FakeImage('vis.tif', [0.5, 0.6]))
    # unequal counts without wavelengths need `force`
    with pytest.raises(ValueError, match='Could not match'):
        match_bands(_FakeImage('p.tif', [None, None]), _FakeImage('q.tif', [None] * 5))
    assert match_bands(_FakeImage('p.tif', [None, None]), _FakeImage('q.tif', [None] * 5), force=True) == \
        ((1, 2), (1, 2))


def _write_pair(tmp_path, ratio=4):
    rng = np.random.default_rng(5)
    ref = rng.integers(20, 200, (4, 30, 26)).astype('uint8')
    src = rng.integers(1, 4000, (3, 30 * ratio - 6, 26 * ratio - 9)).astype('uint16')
    src[:, :3, :] = 0
    ref_tf = Affine(20.0, 0, 1000.0, 0, -20.0, 9000.0)
    src_tf = ref_tf * Affine.scale(1.0 / ratio) * Affine.translation(2, 3)
    src_path = write_geotiff(tmp_path / 'src.tif', src, src_tf, nodata=0, geokeys=GEOKEYS)
    ref_path = write_geotiff(tmp_path / 'ref.tif', ref, ref_tf, geokeys=GEOKEYS,
                             descriptions=['B2', 'B3', 'B4', 'B8'],
                             band_tags=[dict(center_wavelength=w, name=n) for w, n in
                                        zip((0.49, 0.56, 0.665, 0.842), ('B2', 'B3', 'B4', 'B8'))])
    return src, src_tf, src_path, ref, ref_tf, ref_path


def test_file_pair_and_outputs(tmp_path):
    """ RasterFuse / RasterCompare opened from file names: band matching, staged rasters, and the corrected / parameter
    files with the reference's metadata (fuse.py:167-293) -- everything around the GPU call. """
    src, src_tf, src_path, ref, ref_tf, ref_path = _write_pair(tmp_path)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        fuse = RasterFuse(src_path, ref_path)
        compare = RasterCompare(str(src_path), str(ref_path))
    assert fuse.proc_crs == ProcCrs.ref
    assert fuse.src_bands == (1, 2, 3) and fuse.ref_bands == (3, 2, 1)          # RGB assumed -> B4, B3, B2
    assert [compare._band_name(i) for i in range(3)] == ['B4', 'B3', 'B2']
    files = fuse._files
    assert np.array_equal(np.asarray(files.src_ra.array), src) and files.src_ra.nodata == 0
    assert files.src_ra.transform == src_tf
    assert np.array_equal(np.asarray(files.ref_ra.array), ref[[2, 1, 0]]) and np.isnan(files.ref_ra.nodata)
    with pytest.raises(TypeError):
        RasterFuse(src_path, files.ref_ra)

    # outputs, from stand-in results (the GPU call itself is covered by the GPU tier)
    corr = RasterArray(src.astype('float32') * 0.5, files.src_ra.crs, src_tf, nodata=float('nan'))
    params = RasterArray(np.arange(9 * 30 * 26, dtype='float32').reshape(9, 30, 26), files.ref_ra.crs, ref_tf,
                         nodata=float('nan'))
    meta = dict(model=Model.gain_offset, kernel_shape=(5, 5), **RasterFuse.create_model_config(),
                **RasterFuse.create_block_config(threads=2))
    corr_path = tmp_path / ('src' + create_out_postfix(ProcCrs.ref, Model.gain_offset, (5, 5)))
    assert corr_path.name == 'src_FUSE_cREF_mGAIN-OFFSET_k5_5.tif'
    param_path = create_param_filename(corr_path)
    assert param_path.name == 'src_FUSE_cREF_mGAIN-OFFSET_k5_5_PARAM.tif'
    files.write_corrected(corr, corr_path, ProcCrs.ref, RasterFuse.create_out_profile(), **meta)
    files.write_params(params, param_path, ProcCrs.ref, RasterFuse.create_out_profile(), **meta)
    with pytest.raises(FileExistsError):
        files.write_corrected(corr, corr_path, ProcCrs.ref, RasterFuse.create_out_profile(), **meta)
    with GeoTiffReader(corr_path) as im:
        assert im.dtype.name == 'float32' and np.isnan(im.nodata) and im.block_shape == (512, 512)
        assert im.transform == src_tf and im.crs == files.src_ra.crs and im.planar == 2
        assert im.descriptions == ['B4', 'B3', 'B2']
        assert im.tags(1) == dict(center_wavelength='0.665', name='B4')
        tags = im.tags()
        assert tags['FUSE_SRC_FILE'] == 'src.tif' and tags['FUSE_REF_FILE'] == 'ref.tif'
        assert tags['FUSE_PROC_CRS'] == 'ref' and tags['FUSE_MODEL'] == 'gain_offset'
        assert tags['FUSE_KERNEL_SHAPE'] == '(5, 5)' and tags['FUSE_R2_INPAINT_THRESH'] == '0.25'
        assert tags['FUSE_DOWNSAMPLING'] == 'average' and tags['FUSE_UPSAMPLING'] == 'cubic_spline'
        assert tags['FUSE_THREADS'] == '2' and tags['FUSE_MASK_PARTIAL'] == 'False'
        assert np.array_equal(im.read(), corr.array)
    with GeoTiffReader(param_path) as im:
        assert im.count == 9 and im.dtype.name == 'float32' and np.isnan(im.nodata) and im.transform == ref_tf
        assert im.descriptions == ['B4_GAIN', 'B3_GAIN', 'B2_GAIN', 'B4_OFFSET', 'B3_OFFSET', 'B2_OFFSET', 'B4_R2',
                                   'B3_R2', 'B2_R2']
        assert np.array_equal(im.read(), params.array)
    # an existing output is refused before any work is done (fuse.py:275-282)
    with fuse:
        with pytest.raises(FileExistsError):
            fuse.process(corr_path, Model.gain_offset, (5, 5))


@needs_reference
def test_geotiff_reads_gdal_written_files():
    """ The reference's own test images (written by GDAL / Earth Engine): tiles and strips, chunky and band-separate,
    uint8 / int16 / float32, against PIL's decoder and the oracle's minimal reader. """
    from PIL import Image
    from oracle.tiff_min import read_geotiff
    s2 = REF_DATA / 'reference' / 'sentinel2_b432_byte.tif'
    with GeoTiffReader(s2) as im:
        assert np.array_equal(im.read(), np.moveaxis(np.array(Image.open(s2)), 2, 0))
        assert im.colorinterp == ['red', 'green', 'blue'] and im.descriptions == ['B4', 'B3', 'B2']
        assert im.tags(1)['center_wavelength'] == '0.6645' and im.nodata is None
    for rel in ('source/ngi_rgb_byte_3.tif', 'reference/landsat8_byte.tif', 'reference/modis_nbar.tif'):
        with GeoTiffReader(REF_DATA / rel) as im:
            exp = read_geotiff(REF_DATA / rel)
            assert np.array_equal(im.read(), exp['array']) and tuple(im.transform) == exp['transform']
            assert im.nodata == exp['nodata']
    with GeoTiffReader(REF_DATA / 'parameter' / 'float_100cm_rgb_FUSE_cREF_mGAIN-OFFSET_k5_5_PARAM.tif') as im:
        assert im.count == 9 and im.dtype.name == 'float32' and im.block_shape == (16, 16) and np.isnan(im.nodata)
        assert im.descriptions[:4] == ['B1_GAIN', 'B2_GAIN', 'B3_GAIN', 'B1_OFFSET']
        assert im.tags()['FUSE_MODEL'] == 'gain_offset' and im.tags()['FUSE_KERNEL_SHAPE'] == '(5, 5)'
        gain = im.read(1)
        assert np.nanmax(np.abs(gain - 1)) < 1e-5 and np.isnan(gain[0, 0])


@needs_reference
def test_band_matching_of_the_reference_images():
    """ docs/cli.rst:92-114: the matches the reference reports for its own test images. """
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ngi = GeoTiffReader(REF_DATA / 'source' / 'ngi_rgb_byte_1.tif')
        s2 = GeoTiffReader(REF_DATA / 'reference' / 'sentinel2_b432_byte.tif')
        l8 = GeoTiffReader(REF_DATA / 'reference' / 'landsat8_byte.tif')
        modis = GeoTiffReader(REF_DATA / 'reference' / 'modis_nbar.tif')
        assert match_bands(ngi, s2) == ((1, 2, 3), (1, 2, 3))
        assert match_bands(ngi, l8) == ((1, 2, 3), (4, 3, 2))
        assert match_bands(l8, modis, src_bands=[4, 3, 2]) == ((4, 3, 2), (1, 4, 3))
        assert ngi.crs == s2.crs == l8.crs == modis.crs


def test_geotiff_georeferencing_variants(tmp_path):
    """ ModelTransformation instead of scale + tie point; RasterPixelIsPoint (the tie point is a pixel centre); a tie
    point that is not at the raster origin. """
    import struct
    array = np.arange(16 * 16, dtype='uint8').reshape(1, 16, 16)
    path = write_geotiff(tmp_path / 'a.tif', array, TF, geokeys=GEOKEYS, compress=None)
    blob = bytearray(path.read_bytes())
    ifd, = struct.unpack('<I', blob[4:8])
    n, = struct.unpack('<H', blob[ifd:ifd + 2])
    pos = {struct.unpack('<H', blob[ifd + 2 + 12 * i:ifd + 4 + 12 * i])[0]: ifd + 2 + 12 * i for i in range(n)}
    # tie point (i, j) = (2, 3) <-> the same world point
    tie_off, = struct.unpack('<I', blob[pos[33922] + 8:pos[33922] + 12])
    x, y = TF * (2, 3)
    blob[tie_off:tie_off + 48] = struct.pack('<6d', 2.0, 3.0, 0.0, x, y, 0.0)
    (tmp_path / 'tie.tif').write_bytes(bytes(blob))
    with GeoTiffReader(tmp_path / 'tie.tif') as im:
        assert im.transform == TF
    # RasterPixelIsPoint: GTRasterTypeGeoKey (1025) = 2 -> the grid origin is half a pixel up and left of the tie point
    # (the writer itself always writes PixelIsArea -- see test_pixel_is_point_round_trip -- so patch the key in place)
    blob = bytearray(path.read_bytes())
    key_off, = struct.unpack('<I', blob[pos[34735] + 8:pos[34735] + 12])
    keys = list(struct.unpack('<16H', blob[key_off:key_off + 32]))
    assert keys[8:12] == [1025, 0, 1, 1]
    keys[11] = 2
    blob[key_off:key_off + 32] = struct.pack('<16H', *keys)
    (tmp_path / 'point.tif').write_bytes(bytes(blob))
    with GeoTiffReader(tmp_path / 'point.tif') as im:
        assert im.transform == TF * Affine.translation(-0.5, -0.5)
        point_crs, point_keys = im.crs, im.geokeys
    with GeoTiffReader(path) as im:
        assert im.crs == point_crs          # PixelIsPoint / PixelIsArea describe the raster space, not the CRS
    # writing with keys passed through from a PixelIsPoint file must not shift the grid: write -> read -> write -> read
    t0 = TF * Affine.translation(-0.5, -0.5)
    again = write_geotiff(tmp_path / 'again.tif', array, t0, geokeys=point_keys, compress=None)
    with GeoTiffReader(again) as im:
        assert im.transform == t0
        again2 = write_geotiff(tmp_path / 'again2.tif', array, im.transform, crs=im.crs, compress=None)
    with GeoTiffReader(again2) as im:
        assert im.transform == t0
    # ModelTransformation (34264): re-tag the pixel-scale entry as a 16-double matrix and drop the tie point
    blob = bytearray(path.read_bytes())
    matrix = struct.pack('<16d', TF.a, 0, 0, TF.c, 0, TF.e, 0, TF.f, 0, 0, 0, 0, 0, 0, 0, 1)
    blob[pos[33550]:pos[33550] + 12] = struct.pack('<HHII', 34264, 12, 16, len(blob))
    blob[pos[33922]:pos[33922] + 2] = struct.pack('<H', 65000)            # an unknown private tag: ignored
    blob += matrix
    # (the directory is no longer sorted by tag; readers must not rely on the order)
    (tmp_path / 'matrix.tif').write_bytes(bytes(blob))
    with GeoTiffReader(tmp_path / 'matrix.tif') as im:
        assert im.transform == TF and np.array_equal(im.read(), array)


def test_internal_mask_ifd_is_refused(tmp_path):
    """ A GDAL per-dataset mask sits in a second IFD (NewSubfileType bit 2); the reference honours it through
    dataset_mask (raster_array.py:170-197).  It is not decoded here, so the file must be refused, not read as valid. """
    import struct
    array = np.arange(16 * 16, dtype='uint8').reshape(1, 16, 16)
    path = write_geotiff(tmp_path / 'a.tif', array, TF, geokeys=GEOKEYS, compress=None)
    with GeoTiffReader(path) as im:
        assert not im.has_internal_mask
    blob = bytearray(path.read_bytes())
    ifd, = struct.unpack('<I', blob[4:8])
    n, = struct.unpack('<H', blob[ifd:ifd + 2])
    next_pos = ifd + 2 + 12 * n
    assert struct.unpack('<I', blob[next_pos:next_pos + 4])[0] == 0
    # append a second IFD: NewSubfileType = 4 (mask), 16 x 16, 1 bit per sample
    second = len(blob) + (len(blob) & 1)
    blob += b'\0' * (second - len(blob))
    entries = [(254, 4, 1, 4), (256, 3, 1, 16), (257, 3, 1, 16), (258, 3, 1, 1)]
    blob += struct.pack('<H', len(entries))
    for tag, typ, cnt, val in entries:
        blob += struct.pack('<HHII', tag, typ, cnt, val)
    blob += struct.pack('<I', 0)
    blob[next_pos:next_pos + 4] = struct.pack('<I', second)
    (tmp_path / 'masked.tif').write_bytes(bytes(blob))
    with pytest.raises(NotImplementedError, match='mask'):
        GeoTiffReader(tmp_path / 'masked.tif')
    # a .msk side-car is the other form of the same thing
    (tmp_path / 'a.tif.msk').write_bytes(b'')
    with pytest.raises(NotImplementedError, match='mask'):
        GeoTiffReader(path)


def test_band_stream_writer_equals_eager_writer(tmp_path):
    """ write_geotiff fed band by band (geotiff.BandStream: the file path's overlap of encoding with the GPU step,
    SURVEY.md 8f-4) writes the same bytes as the eager call; a failing producer surfaces as an error, not a hang. """
    import threading
    import time
    from homonim_b200.geometry import Affine
    from homonim_b200.geotiff import BandStream, GeoTiffReader, write_geotiff
    rng = np.random.default_rng(8)
    data = rng.integers(0, 4000, (3, 300, 420)).astype('uint16')
    tf = Affine(2.0, 0, 100.0, 0, -2.0, 900.0)
    for compress in ('deflate', None):
        eager = write_geotiff(tmp_path / f'eager_{compress}.tif', data, tf, nodata=0, compress=compress, blocksize=128)
        buf = np.zeros_like(data)
        stream = BandStream(buf)

        def produce():
            for b in (0, 1, 2):
                time.sleep(0.05)
                buf[b] = data[b]
                stream.set_ready(b)
        t = threading.Thread(target=produce)
        t.start()
        lazy = write_geotiff(tmp_path / f'lazy_{compress}.tif', stream, tf, nodata=0, compress=compress, blocksize=128)
        t.join()
        assert pathlib.Path(eager).read_bytes() == pathlib.Path(lazy).read_bytes()
        with GeoTiffReader(lazy) as im:
            assert np.array_equal(im.read(), data)
    # pixel-interleaved output needs every band before the first tile: still correct
    buf = np.zeros_like(data)
    stream = BandStream(buf)
    threading.Thread(target=lambda: [(buf.__setitem__(b, data[b]), stream.set_ready(b)) for b in range(3)]).start()
    p = write_geotiff(tmp_path / 'pixel.tif', stream, tf, interleave='pixel', blocksize=128)
    with GeoTiffReader(p) as im:
        assert np.array_equal(im.read(), data)
    # a producer that fails
    stream = BandStream(np.zeros_like(data))
    threading.Timer(0.05, lambda: stream.fail(ValueError('boom'))).start()
    with pytest.raises(RuntimeError, match='producer'):
        write_geotiff(tmp_path / 'fail.tif', stream, tf, blocksize=128)
    assert not (tmp_path / 'fail.tif').exists() and not (tmp_path / 'fail.tif.part').exists()

"""
oracle/tiff_min.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A minimal baseline-TIFF / GeoTIFF reader (classic TIFF, uncompressed or deflate, predictor 1 or 2, strips or tiles,
chunky or planar, 8 / 16-bit integer samples, north-up ModelPixelScale + ModelTiepoint georeferencing).  It exists only
so that the fixture generators can read the reference's own test images (/root/reference/tests/data) in the build
container, where neither rasterio nor GDAL is installed.
"""
import re
import struct
import zlib

import numpy as np

_TYPES = {1: ('B', 1), 2: ('c', 1), 3: ('H', 2), 4: ('I', 4), 5: ('II', 8), 6: ('b', 1), 7: ('B', 1), 8: ('h', 2),
          9: ('i', 4), 11: ('f', 4), 12: ('d', 8)}


def read_tags(path):
    """ (file bytes, byte-order char, {tag: value tuple or str}) of the first IFD. """
    buf = open(path, 'rb').read()
    bo = '<' if buf[:2] == b'II' else '>'
    if struct.unpack(bo + 'H', buf[2:4])[0] != 42:
        raise ValueError(f'{path}: not a classic TIFF')
    off, = struct.unpack(bo + 'I', buf[4:8])
    n, = struct.unpack(bo + 'H', buf[off:off + 2])
    tags = {}
    for i in range(n):
        entry = buf[off + 2 + 12 * i:off + 14 + 12 * i]
        tag, typ, cnt = struct.unpack(bo + 'HHI', entry[:8])
        fmt, size = _TYPES[typ]
        total = size * cnt
        data = entry[8:8 + total] if total <= 4 else buf[struct.unpack(bo + 'I', entry[8:12])[0]:][:total]
        if typ == 2:
            tags[tag] = data.decode('latin1').rstrip('\x00')
        elif typ == 5:
            tags[tag] = struct.unpack(bo + 'I' * (2 * cnt), data)
        else:
            tags[tag] = struct.unpack(bo + fmt * cnt, data)
    return buf, bo, tags


def read_geotiff(path):
    """
    Returns dict(array=[bands, height, width], transform=(a, b, c, d, e, f), nodata=float or None,
    descriptions=[str or None per band], wavelengths=[float or None per band]).
    """
    buf, bo, t = read_tags(path)
    width, height = t[256][0], t[257][0]
    bits, spp = t[258], t.get(277, (1,))[0]
    if len(set(bits)) != 1 or bits[0] not in (8, 16):
        raise ValueError(f'{path}: unsupported BitsPerSample {bits}')
    fmt = t.get(339, (1,))[0]
    dtype = np.dtype({(8, 1): 'u1', (16, 1): 'u2', (16, 2): 'i2', (8, 2): 'i1'}[(bits[0], fmt)]).newbyteorder(bo)
    compression, predictor, planar = t.get(259, (1,))[0], t.get(317, (1,))[0], t.get(284, (1,))[0]
    if compression not in (1, 8, 32946):
        raise ValueError(f'{path}: unsupported compression {compression}')
    tiled = 322 in t
    if tiled:
        bw, bh = t[322][0], t[323][0]
        offsets, counts = t[324], t[325]
    else:
        bw, bh = width, t.get(278, (height,))[0]
        offsets, counts = t[273], t[279]
    across, down = -(-width // bw), -(-height // bh)
    chunk_spp = spp if planar == 1 else 1
    planes = 1 if planar == 1 else spp
    if len(offsets) != across * down * planes:
        raise ValueError(f'{path}: unexpected block count {len(offsets)}')
    out = np.zeros((spp, height, width), dtype=dtype.newbyteorder('='))
    for plane in range(planes):
        for by in range(down):
            for bx in range(across):
                idx = (plane * down + by) * across + bx
                raw = buf[offsets[idx]:offsets[idx] + counts[idx]]
                if compression != 1:
                    raw = zlib.decompress(raw)
                rows = bh if tiled else min(bh, height - by * bh)
                block = np.frombuffer(raw, dtype=dtype, count=rows * bw * chunk_spp).reshape(rows, bw, chunk_spp)
                if predictor == 2:
                    block = np.cumsum(block, axis=1, dtype=dtype)
                elif predictor != 1:
                    raise ValueError(f'{path}: unsupported predictor {predictor}')
                y0, x0 = by * bh, bx * bw
                hh, ww = min(rows, height - y0), min(bw, width - x0)
                if planar == 1:
                    out[:, y0:y0 + hh, x0:x0 + ww] = np.moveaxis(block[:hh, :ww, :], 2, 0)
                else:
                    out[plane, y0:y0 + hh, x0:x0 + ww] = block[:hh, :ww, 0]
    scale, tie = t[33550], t[33922]
    if tie[0] != 0 or tie[1] != 0:
        raise ValueError(f'{path}: tie point is not at the raster origin')
    transform = (float(scale[0]), 0.0, float(tie[3]), 0.0, -float(scale[1]), float(tie[4]))
    nodata = float(t[42113]) if 42113 in t else None
    meta = t.get(42112, '')
    descriptions, wavelengths = [None] * spp, [None] * spp
    for m in re.finditer(r'<Item name="DESCRIPTION" sample="(\d+)" role="description">([^<]*)</Item>', meta):
        descriptions[int(m.group(1))] = m.group(2)
    for m in re.finditer(r'<Item name="center_wavelength" sample="(\d+)"[^>]*>([^<]*)</Item>', meta):
        wavelengths[int(m.group(1))] = float(m.group(2))
    return dict(array=out, transform=transform, nodata=nodata, descriptions=descriptions, wavelengths=wavelengths,
                photometric=t.get(262, (None,))[0])

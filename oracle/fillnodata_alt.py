"""
oracle/fillnodata_alt.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A second, independent restatement of ``rasterio.fill.fillnodata(image, mask, max_search_distance=100,
smoothing_iterations=0)`` == ``GDALFillNodata`` (alg/rasterfill.cpp), written from the algorithm's description rather
than from oracle/gdal_restate.c, to catch transcription errors in the latter (GDAL itself cannot be installed in this
image: in-painting parity stays UNPINNED to GDAL).  Brute force, pure numpy / Python: for small
rasters only.

The algorithm, as described:
  * pixels with mask != 0 are sources, pixels with mask == 0 are filled; only original sources are ever used;
  * for a pixel (y, x) to fill, the columns x - s (left) and x + s (right), s = 0 .. floor(max_search_distance), are
    examined (positions beyond the raster are clamped to its first / last column).  In a column, the nearest source at
    or above row y is a candidate for the TOP quadrant of that side, the nearest source at or below row y for the
    BOTTOM quadrant; the centre column (s = 0) belongs to the left side only;
  * a quadrant keeps the candidate with the smallest Euclidean distance (first come wins ties, columns nearest first);
  * value = sum(v_q / d_q) / sum(1 / d_q) over the quadrants whose distance is <= max_search_distance, accumulated in
    double and stored as float32; a pixel for which no quadrant found a source is left unchanged.
Consequences worth testing: a source on the pixel's own ROW is found by the top and the bottom quadrant of its side (it
counts twice); a source in the pixel's own COLUMN counts once (left side only).
"""
import math

import numpy as np


def fillnodata_alt(image, mask, max_search_distance=100.0):
    image = np.asarray(image, dtype='float32')
    mask = np.asarray(mask) != 0
    h, w = image.shape
    out = image.copy()
    radius = int(math.floor(max_search_distance))
    rows_of = [np.flatnonzero(mask[:, x]) for x in range(w)]
    for y, x in zip(*np.nonzero(~mask)):
        # quadrant -> [distance, value]; order of the final sum: top-left, bottom-left, top-right, bottom-right
        best = {q: [float(max_search_distance) + 1.0, 0.0] for q in ('tl', 'bl', 'tr', 'br')}
        for s in range(radius + 1):
            for side in ('l', 'r'):
                if side == 'r' and s == 0:
                    continue
                cx = min(max(x - s if side == 'l' else x + s, 0), w - 1)
                rows = rows_of[cx]
                if rows.size == 0:
                    continue
                above = np.searchsorted(rows, y, side='right') - 1          # last source row <= y
                below = np.searchsorted(rows, y, side='left')               # first source row >= y
                for quad, idx in (('t' + side, above), ('b' + side, below)):
                    if idx < 0 or idx >= rows.size:
                        continue
                    sy = int(rows[idx])
                    d2 = float((cx - x) ** 2 + (sy - y) ** 2)
                    if d2 < best[quad][0] * best[quad][0]:
                        best[quad] = [math.sqrt(d2), float(image[sy, cx])]
        wsum = vsum = 0.0
        for quad in ('tl', 'bl', 'tr', 'br'):
            d, v = best[quad]
            if d <= max_search_distance:
                wsum += 1.0 / d
                vsum += v * (1.0 / d)
        if wsum > 0.0:
            out[y, x] = np.float32(vsum / wsum)
    return out

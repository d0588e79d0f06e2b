"""
Source <-> reference band matching for rasters opened from files (reference homonim/matched_pair.py:95-342): which
bands of each file take part, and which reference band corrects which source band.  Same rules as the reference:

* alpha bands and ``*_MASK`` / ``*_DIST`` helper bands are never used; bands carrying ``center_wavelength`` metadata
  are preferred; a three-band image without wavelengths is taken to be RGB (standard red / green / blue wavelengths,
  by colour interpretation where the file has one, else in file order);
* when both images have wavelengths, source bands are paired greedily with the reference band of the nearest centre
  wavelength (relative distance, each reference band used once); a pairing further apart than 10 % is an error;
* bands left without wavelengths are paired in file order when the counts agree (or when ``force`` is set).
"""
import logging
import math
import pathlib
import warnings
from typing import List, Optional, Sequence, Tuple

from homonim_b200.errors import BandMatchWarning

logger = logging.getLogger(__name__)

MAX_REL_WAVELENGTH_DIFF = 0.1        # matched_pair.py:36
STANDARD_RGB = dict(red=0.650, green=0.560, blue=0.480)          # matched_pair.py:152-154


def _short(name: str) -> str:
    return pathlib.Path(str(name)).name


def band_info(im, bands: Optional[Sequence[int]] = None) -> Tuple[List[int], List[str], List[float]]:
    """
    The bands of ``im`` to use (1-based), their names and centre wavelengths (NaN where unknown) --
    ``MatchedPairReader._get_band_info``, matched_pair.py:95-180.  ``im`` is a
    :class:`~homonim_b200.geotiff.GeoTiffReader` (or anything with ``count``, ``colorinterp``, ``descriptions``,
    ``tags(bidx)`` and ``name``).
    """
    name = _short(getattr(im, 'name', ''))

    def helper_band(i):          # geedim mask / distance bands
        descr = im.descriptions[i]
        return bool(descr) and (descr.endswith('_MASK') or descr.endswith('_DIST'))

    usable = [i + 1 for i in range(im.count) if im.colorinterp[i] != 'alpha' and not helper_band(i)]
    with_wavelength = [b for b in usable if 'center_wavelength' in im.tags(b)]
    if bands is not None and len(bands) > 0:
        bands = [int(b) for b in bands]
        invalid = sorted(set(bands) - set(range(1, im.count + 1)))
        if invalid:
            raise ValueError(f'User specified {name} bands contain invalid band(s) {invalid}.')
        alpha = sorted(set(bands) - set(usable))
        if alpha:
            raise ValueError(f'User specified {name} bands contain alpha band(s) {alpha}.')
        if with_wavelength and not set(bands) <= set(with_wavelength):
            warnings.warn(f'User specified {name} bands contain non-reflectance band index(es) '
                          f'{sorted(set(bands) - set(with_wavelength))}.', category=BandMatchWarning)
        chosen = bands
    elif with_wavelength:
        chosen = with_wavelength
    elif usable:
        chosen = usable
    else:
        raise ValueError(f'There are no non-alpha/reflectance {name} bands to use.')

    wavelengths = [float(im.tags(b)['center_wavelength']) if 'center_wavelength' in im.tags(b) else math.nan
                   for b in range(1, im.count + 1)]
    if len(usable) == 3:
        assigned = []
        for b in usable:
            interp = im.colorinterp[b - 1]
            if math.isnan(wavelengths[b - 1]) and interp in STANDARD_RGB:
                wavelengths[b - 1] = STANDARD_RGB[interp]
                assigned.append(interp)
        if assigned:
            warnings.warn(f'Assigning standard {", ".join(assigned)} center wavelengths for {name}.',
                          category=BandMatchWarning)
        if all(math.isnan(wavelengths[b - 1]) for b in usable):
            warnings.warn(f'Assuming image is RGB, and assigning standard center wavelengths for {name}.',
                          category=BandMatchWarning)
            for b, w in zip(usable, STANDARD_RGB.values()):
                wavelengths[b - 1] = w
    names = [im.descriptions[b - 1] or str(b) for b in chosen]
    return list(chosen), names, [wavelengths[b - 1] for b in chosen]


def _greedy_match(dist: List[List[float]]) -> Tuple[List[Optional[int]], List[float]]:
    """ Repeatedly pair the (row, column) with the smallest distance still available; every row and column is used at
    most once (matched_pair.py:249-279).  Returns per row the matched column (or None) and its distance. """
    n_rows, n_cols = len(dist), (len(dist[0]) if dist else 0)
    match: List[Optional[int]] = [None] * n_rows
    match_dist = [math.nan] * n_rows
    free_rows, free_cols = set(range(n_rows)), set(range(n_cols))
    while True:
        best = None
        for r in sorted(free_rows):
            for c in sorted(free_cols):
                d = dist[r][c]
                if not math.isnan(d) and (best is None or d < best[0]):
                    best = (d, r, c)
        if best is None:
            break
        d, r, c = best
        match[r], match_dist[r] = c, d
        free_rows.discard(r)
        free_cols.discard(c)
    return match, match_dist


def match_bands(src_im, ref_im, src_bands: Optional[Sequence[int]] = None, ref_bands: Optional[Sequence[int]] = None,
                force: bool = False) -> Tuple[Tuple[int, ...], Tuple[int, ...]]:
    """ Matching (source bands, reference bands), 1-based -- ``MatchedPairReader._match_pair_bands``,
    matched_pair.py:228-342. """
    s_bands, s_names, s_wl = band_info(src_im, src_bands)
    r_bands, r_names, r_wl = band_info(ref_im, ref_bands)
    src_name, ref_name = _short(getattr(src_im, 'name', 'source')), _short(getattr(ref_im, 'name', 'reference'))
    if len(s_bands) > len(r_bands):
        if not force:
            raise ValueError(f'{ref_name} has fewer bands than {src_name}.')
        warnings.warn(f'{ref_name} has fewer bands than {src_name}.', category=BandMatchWarning)

    matched: List[Optional[int]] = [None] * len(s_bands)          # reference band per source band
    # (the reference tests `any(wavelengths)` on arrays in which unknown wavelengths are NaN, i.e. truthy: the
    # wavelength pass runs whenever it is not forced off, and pairs only bands that do have wavelengths)
    if s_wl and r_wl and not force:
        rel = [[abs(sw - rw) / sw if not (math.isnan(sw) or math.isnan(rw)) else math.nan for rw in r_wl]
               for sw in s_wl]
        idx, dist = _greedy_match(rel)
        too_far = [i for i, d in enumerate(dist) if not math.isnan(d) and d > MAX_REL_WAVELENGTH_DIFF]
        if too_far:
            raise ValueError(
                f'{src_name} band(s) {[s_names[i] for i in too_far]} could not be auto-matched.  The nearest '
                f'{ref_name} band(s) were {[r_names[idx[i]] for i in too_far]}, at center wavelength difference(s) of '
                f'{[round(dist[i], 3) for i in too_far]} (um) respectively.'
            )
        for i, c in enumerate(idx):
            if c is not None:
                matched[i] = r_bands[c]
        logger.debug(f'Matched {src_name} band(s) {[s_names[i] for i, c in enumerate(idx) if c is not None]} by '
                     f'wavelength with {ref_name} band(s) {[r_names[c] for c in idx if c is not None]}.')

    n_matched = sum(m is not None for m in matched)
    if n_matched < min(len(s_bands), len(r_bands)):
        open_src = [i for i, m in enumerate(matched) if m is None]
        open_ref = [b for b in r_bands if b not in matched]
        if len(s_bands) == len(r_bands):
            for i, b in zip(open_src, open_ref):                   # file order
                matched[i] = b
        elif force:
            for i, b in zip(open_src, open_ref):
                matched[i] = b
        else:
            raise ValueError(
                f'Could not match {src_name} band(s) {[s_bands[i] for i in open_src]} with {ref_name} band(s) '
                f'{open_ref}.  Ensure {src_name} and {ref_name} non-alpha band counts match, {src_name} and '
                f'{ref_name} have ``center_wavelength`` tags for each band, or set `force` to True.'
            )
    pairs = [(s, m) for s, m in zip(s_bands, matched) if m is not None]
    return tuple(p[0] for p in pairs), tuple(p[1] for p in pairs)

"""
oracle/kernel_model_np.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy + OpenCV restatement ("port") of the reference's per-pixel kernel-model fit / apply path,
/root/reference/homonim/kernel_model.py (v0.4.3), on plain arrays + 6-tuple geo-transforms.  Every function cites
the reference lines it follows.  It keeps the reference's precision class for every intermediate as executed with
numpy >= 2 (NEP 50 promotion) and cv2 4.x, SURVEY.md section 8(a) "numerics note":

  * cv.boxFilter(f32) -> f32 (double accumulation inside OpenCV), cv.sqrBoxFilter(f32) -> f64;
  * the gain-offset numerator is all-float32, the denominator float64 - float32;
  * gain-blk-offset normalises the source with float64 scalars, so its window sums are float64.

Pinning: tests/test_oracle_golden.py checks this module BIT-FOR-BIT against tests/golden/*.npz, which
oracle/make_golden.py produced by running the UNMODIFIED reference (through oracle/rasterio_stub) in the build
container.  The GDAL-backed steps (re-projection, fillnodata) go through oracle/gdal_restate.py in both cases and are
PARITY UNPINNED against GDAL itself (GDAL is not installed here).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from typing import Optional, Sequence, Tuple

import cv2 as cv
import numpy as np

from oracle import gdal_restate

NODATA = float('nan')      # RasterArray.default_nodata, raster_array.py:48
F32 = 'float32'            # RasterArray.default_dtype, raster_array.py:49
_BOX = dict(normalize=False, borderType=cv.BORDER_CONSTANT)   # kernel_model.py:155, 256, 331

MODEL_GAIN = 'gain'
MODEL_GAIN_BLK_OFFSET = 'gain-blk-offset'
MODEL_GAIN_OFFSET = 'gain-offset'


# ---------------------------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------------------------
def nan_equals(a, b):
    """ utils.py:54-56 """
    return (a == b) | (np.isnan(a) & np.isnan(b))


def valid_mask(array: np.ndarray, nodata) -> np.ndarray:
    """ RasterArray.mask, raster_array.py:298-308 (2D mask; a 3D pixel is valid if any band is). """
    if nodata is None:
        return np.full(array.shape[-2:], True)
    mask = ~nan_equals(array, nodata)
    if array.ndim > 2:
        mask = np.any(mask, axis=0)
    return mask


def as_working(array: np.ndarray) -> np.ndarray:
    """ Integer rasters are read as float32 (raster_array.py:178-188, out_dtype=float32); floats are kept. """
    array = np.asarray(array)
    return array if np.issubdtype(array.dtype, np.floating) else array.astype(F32)


def _ksize(kernel_shape) -> Tuple[int, int]:
    """ OpenCV wants (width, height): kernel_shape[::-1], kernel_model.py:257. """
    return (int(kernel_shape[1]), int(kernel_shape[0]))


def _box(x, kernel_shape):
    return cv.boxFilter(x, -1, _ksize(kernel_shape), **_BOX)


def _sqr_box(x, kernel_shape):
    return cv.sqrBoxFilter(x, -1, _ksize(kernel_shape), **_BOX)


def res_of(transform) -> Tuple[float, float]:
    """ RasterArray.res, raster_array.py:271-274 """
    return float(transform[0]), -float(transform[4])


def pick_resampling(from_res, to_res, downsampling='average', upsampling='cubic_spline') -> str:
    """ KernelModel._get_resampling, kernel_model.py:138-140 """
    return downsampling if np.prod(np.abs(from_res)) <= np.prod(np.abs(to_res)) else upsampling


# ---------------------------------------------------------------------------------------------------------------------
# same-grid fit (kernel_model.py:142-373)
# ---------------------------------------------------------------------------------------------------------------------
def r2_plane(ref, src, params, mask, kernel_shape, mask_sum=None, ref_sum=None, src_sum=None, ref2_sum=None,
             src2_sum=None, src_ref_sum=None, dest=None):
    """ KernelModel._r2_array, kernel_model.py:142-214 (mask already applied to ref / src). """
    if mask_sum is None:
        mask_sum = _box(mask.astype(F32), kernel_shape)                      # :167
    if ref_sum is None:
        ref_sum = _box(ref, kernel_shape)                                    # :169
    if ref2_sum is None:
        ref2_sum = _sqr_box(ref, kernel_shape)                               # :171
    if src2_sum is None:
        src2_sum = _sqr_box(src, kernel_shape)                               # :173
    if src_ref_sum is None:
        src_ref_sum = _box(src * ref, kernel_shape)                          # :175

    ss_tot = (mask_sum * ref2_sum) - (ref_sum ** 2)                          # :179
    if params.shape[0] > 1:
        if src_sum is None:
            src_sum = _box(src, kernel_shape)                                # :184
        ss_res = (
            ((params[0] ** 2) * src2_sum) +
            (2 * np.prod(params[:2], axis=0) * src_sum) -
            (2 * params[0] * src_ref_sum) -
            (2 * params[1] * ref_sum) +
            ref2_sum + (mask_sum * (params[1] ** 2))
        )                                                                    # :189-195
    else:
        ss_res = (((params[0] ** 2) * src2_sum) - (2 * params[0] * src_ref_sum) + ref2_sum)   # :201
    ss_res *= mask_sum                                                       # :203
    if dest is None:
        dest = np.full(src.shape, NODATA, dtype=F32)                         # :207-209
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        np.divide(ss_res, ss_tot, out=dest, where=mask)                      # :212
        np.subtract(1, dest, out=dest, where=mask)                           # :213
    return dest


def block_norm(src, ref, mask) -> np.ndarray:
    """ KernelModel._fit_block_norm, kernel_model.py:216-229 """
    norm = np.zeros(2)
    if not np.any(mask):
        return norm
    norm[0] = np.std(ref[mask]) / np.std(src[mask])
    norm[1] = np.percentile(ref[mask], 1) - np.percentile(src[mask], 1) * norm[0]
    return norm


def _fit_gain_core(src, ref, mask, kernel_shape, find_r2):
    """ KernelModel._fit_gain, kernel_model.py:231-274; src / ref are private copies, mask = valid in both. """
    ref[~mask] = 0                                                           # :246
    src[~mask] = 0                                                           # :247
    src_sum = _box(src, kernel_shape)                                        # :257
    ref_sum = _box(ref, kernel_shape)                                        # :258
    params = np.full((3 if find_r2 else 2, *src.shape), NODATA, dtype=F32)   # :261
    params[1, mask] = 0                                                      # :262
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        np.divide(ref_sum, src_sum, out=params[0], where=mask)               # :265
    if find_r2:
        r2_plane(ref, src, params[:1], mask, kernel_shape, ref_sum=ref_sum, src_sum=src_sum, dest=params[2])   # :269
    return params


def fit_gain(src, src_nodata, ref, ref_nodata, kernel_shape, find_r2=False):
    src, ref = as_working(src).copy(), as_working(ref).copy()
    mask = valid_mask(ref, ref_nodata) & valid_mask(src, src_nodata)         # :245
    return _fit_gain_core(src, ref, mask, kernel_shape, find_r2)


def fit_gain_blk_offset(src, src_nodata, ref, ref_nodata, kernel_shape, find_r2=False):
    """ KernelModel._fit_gain_blk_offset, kernel_model.py:276-303 """
    src, ref = as_working(src).copy(), as_working(ref).copy()
    src_mask = valid_mask(src, src_nodata)
    ref_mask = valid_mask(ref, ref_nodata)
    norm = block_norm(src, ref, ref_mask & src_mask)                         # :289
    # :292 -- the nodata setter (raster_array.py:334-351) rewrites invalid pixels to nan
    if src_nodata is not None and not nan_equals(NODATA, src_nodata):
        src[~src_mask] = NODATA
    # (nodata None: the setter just relabels nodata as nan; any nan already in the data becomes invalid)
    src = (src * norm[0]) + norm[1]                                          # :295 -- float64 under numpy >= 2
    mask = ref_mask & valid_mask(src, NODATA)
    params = _fit_gain_core(src, ref, mask, kernel_shape, find_r2)           # :298
    params[1] = params[0] * norm[1]                                          # :301
    params[0] *= norm[0]                                                     # :302
    return params


def fit_gain_offset(src, src_nodata, ref, ref_nodata, kernel_shape, find_r2=False, r2_inpaint_thresh=0.25):
    """ KernelModel._fit_gain_offset, kernel_model.py:305-373 """
    src, ref = as_working(src).copy(), as_working(ref).copy()
    mask = valid_mask(ref, ref_nodata) & valid_mask(src, src_nodata)         # :319
    ref[~mask] = 0
    src[~mask] = 0
    want_r2 = find_r2 or (r2_inpaint_thresh is not None)                     # :325
    src_sum = _box(src, kernel_shape)                                        # :332
    ref_sum = _box(ref, kernel_shape)                                        # :333
    src_ref_sum = _box(src * ref, kernel_shape)                              # :334
    mask_sum = _box(mask.astype(F32, copy=False), kernel_shape)              # :335-337
    num = (mask_sum * src_ref_sum) - (src_sum * ref_sum)                     # :338
    src2_sum = _sqr_box(src, kernel_shape)                                   # :341
    den = (mask_sum * src2_sum) - (src_sum ** 2)                             # :342
    params = np.full((3 if want_r2 else 2, *src.shape), NODATA, dtype=F32)   # :345
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        np.divide(num, den, out=params[0], where=mask)                       # :348
        np.divide(ref_sum - (params[0] * src_sum), mask_sum, out=params[1], where=mask)   # :351
    if want_r2:
        r2_plane(ref, src, params[:2], mask, kernel_shape, mask_sum=mask_sum, ref_sum=ref_sum, src_sum=src_sum,
                 src2_sum=src2_sum, src_ref_sum=src_ref_sum, dest=params[2])                # :355-359
    if r2_inpaint_thresh is not None:
        with np.errstate(invalid='ignore'):
            r2_mask = (params[2] > r2_inpaint_thresh) & (params[0] > 0) & mask              # :363
        params[1] = gdal_restate.fillnodata(params[1], r2_mask)              # :366
        params[:, ~mask] = NODATA                                            # :367
        r2_mask = ~r2_mask & mask                                            # :370
        with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
            np.divide(ref_sum - mask_sum * params[1], src_sum, out=params[0], where=r2_mask)   # :371
    return params


def fit_same_grid(src, src_nodata, ref, ref_nodata, model, kernel_shape, find_r2=False, r2_inpaint_thresh=0.25):
    """ KernelModel.fit, kernel_model.py:411-440 (inputs are NOT mutated here). """
    if np.shape(src) != np.shape(ref):
        raise ValueError("'ref_ra' and 'src_ra' must have the same CRS, transform and shape")   # :430-431
    model = getattr(model, 'value', model)
    if model == MODEL_GAIN:
        return fit_gain(src, src_nodata, ref, ref_nodata, kernel_shape, find_r2)
    elif model == MODEL_GAIN_BLK_OFFSET:
        return fit_gain_blk_offset(src, src_nodata, ref, ref_nodata, kernel_shape, find_r2)
    return fit_gain_offset(src, src_nodata, ref, ref_nodata, kernel_shape, find_r2, r2_inpaint_thresh)


def apply_same_grid(src, params):
    """ KernelModel.apply, kernel_model.py:442-463 """
    if np.shape(src) != np.shape(params)[-2:]:
        raise ValueError("'param_ra' and 'src_ra' must have the same CRS, transform and shape")   # :459-460
    return (params[0] * as_working(src)) + params[1]                         # :461


def convert_dtype(corr, dtype, nodata=None):
    """
    RasterArray._convert_array_dtype, raster_array.py:353-387, for a float32 corrected plane with NaN nodata (what
    to_rio_dataset converts before writing, :493-500): promote (:366), round half to even when going to an integer
    type (:369-370), clip to its range (:373-377), cast (:380-381), nodata where the plane was NaN (:384-385).
    """
    corr = np.asarray(corr, dtype='float32')
    mask = ~np.isnan(corr)
    to_int = np.issubdtype(np.dtype(dtype), np.integer)
    array = corr.astype(np.promote_types('float32', dtype), copy=True)       # :366
    if to_int:
        np.round(array, out=array)                                           # :370
        info = np.iinfo(dtype)
        np.clip(array, info.min, info.max, out=array)                        # :377
    with np.errstate(invalid='ignore', over='ignore'):
        array = array.astype(dtype, copy=False, casting='unsafe')            # :381
    if nodata is not None:
        array[~mask] = nodata                                                # :385
    return array


# ---------------------------------------------------------------------------------------------------------------------
# RasterCompare accuracy sums and statistics (compare.py:142-187, 232-256) -- SURVEY.md 8f-3
# ---------------------------------------------------------------------------------------------------------------------
COMPARE_SUM_KEYS = ('src_sum', 'ref_sum', 'src2_sum', 'ref2_sum', 'src_ref_sum', 'res2_sum', 'mask_sum')


def compare_sums(src, src_nodata, ref, ref_nodata, dtype='float32') -> dict:
    """
    get_block_sums, compare.py:241-254, for two planes on one grid.  ``dtype='float32'`` is what the reference does
    (float32 terms, numpy's pairwise float32 summation); ``dtype='float64'`` sums the SAME float32 terms in double --
    the sums the CUDA path produces.
    """
    src, ref = as_working(np.array(src, copy=True)), as_working(np.array(ref, copy=True))
    mask = valid_mask(ref, ref_nodata) & valid_mask(src, src_nodata)         # :245
    src[~mask] = 0                                                           # :246
    ref[~mask] = 0                                                           # :247
    terms = dict(src_sum=src, ref_sum=ref, src2_sum=src ** 2, ref2_sum=ref ** 2, src_ref_sum=src * ref,
                 res2_sum=(ref - src) ** 2)                                  # :249-253 (float32 terms)
    sums = {k: v.sum(dtype=dtype) for k, v in terms.items()}
    sums['mask_sum'] = mask.sum()
    return sums


def compare_band_stats(src_sum=0, ref_sum=0, src2_sum=0, ref2_sum=0, src_ref_sum=0, res2_sum=0, mask_sum=0) -> dict:
    """ get_band_stats, compare.py:145-163: Pearson's r squared, RMSE and RMSE relative to the reference mean. """
    src_mean = src_sum / mask_sum                                            # :152
    ref_mean = ref_sum / mask_sum                                            # :153
    pcc_num = src_ref_sum - (mask_sum * src_mean * ref_mean)                 # :154
    pcc_den = (np.sqrt(src2_sum - (mask_sum * (src_mean ** 2))) *
               np.sqrt(ref2_sum - (mask_sum * (ref_mean ** 2))))             # :155-157
    pcc = pcc_num / pcc_den                                                  # :158
    rmse = np.sqrt(res2_sum / mask_sum)                                      # :161
    rrmse = rmse / ref_mean                                                  # :162
    return dict(r2=pcc ** 2, rmse=rmse, rrmse=rrmse, n=int(mask_sum))        # :163


def compare_image_stats(image_sums, band_names) -> dict:
    """ _get_image_stats, compare.py:165-187: per-band statistics keyed by band name, plus their 'Mean' over bands. """
    image_stats, sum_over_bands = {}, {}
    for name, band_sums in zip(band_names, image_sums):
        band_stats = compare_band_stats(**band_sums)
        image_stats[name] = band_stats                                       # :176
        sum_over_bands = {k: sum_over_bands.get(k, 0) + v for k, v in band_stats.items()}       # :177
    image_stats['Mean'] = {k: int(v / len(image_sums)) if isinstance(v, int) else (v / len(image_sums))
                           for k, v in sum_over_bands.items()}               # :180-185
    return image_stats


# ---------------------------------------------------------------------------------------------------------------------
# grid-changing wrappers (kernel_model.py:375-409, 466-535)
# ---------------------------------------------------------------------------------------------------------------------
def _set_mask(array, mask, nodata=NODATA):
    """ RasterArray.mask setter, raster_array.py:310-318 """
    if array.ndim == 2:
        array[~mask] = nodata
    else:
        array[:, ~mask] = nodata


def full_coverage_mask(in_mask, in_transform, params2, param_transform, kernel_shape):
    """ KernelModel._full_coverage_mask, kernel_model.py:375-409; returns a uint8 mask on the param grid. """
    frac = gdal_restate.reproject_array(in_mask.astype('uint8'), in_transform, None, params2.shape[-2:],
                                        param_transform, None, 'average')    # :397
    mask = (frac >= 1).astype('uint8', copy=False)                           # :399
    mask &= valid_mask(params2, NODATA)                                      # :401
    se = cv.getStructuringElement(cv.MORPH_RECT, (int(kernel_shape[1]) + 2, int(kernel_shape[0]) + 2))   # :407
    return cv.erode(mask, se, borderType=cv.BORDER_CONSTANT, borderValue=0)  # :408


def refspace_fit(src, src_transform, src_nodata, ref, ref_transform, ref_nodata, model, kernel_shape, find_r2=False,
                 r2_inpaint_thresh=0.25, downsampling='average', upsampling='cubic_spline'):
    """ RefSpaceModel.fit, kernel_model.py:476-482; returns params on the reference grid. """
    resampling = pick_resampling(res_of(src_transform), res_of(ref_transform), downsampling, upsampling)   # :478
    src_ds = gdal_restate.reproject_array(src, src_transform, src_nodata, np.shape(ref)[-2:], ref_transform, NODATA,
                                          resampling)                        # :480
    return fit_same_grid(src_ds, NODATA, ref, ref_nodata, model, kernel_shape, find_r2, r2_inpaint_thresh)     # :482


def refspace_apply(src, src_transform, src_nodata, params, param_transform, kernel_shape, mask_partial=False,
                   downsampling='average', upsampling='cubic_spline'):
    """ RefSpaceModel.apply, kernel_model.py:484-503; returns the corrected array on the source grid. """
    src_w = as_working(src)
    params2 = np.asarray(params)[:2]                                         # :487
    resampling = pick_resampling(res_of(param_transform), res_of(src_transform), downsampling, upsampling)   # :489
    param_us = gdal_restate.reproject_array(params2, param_transform, NODATA, src_w.shape, src_transform, NODATA,
                                            resampling)                      # :491
    src_mask = valid_mask(src_w, src_nodata)
    if mask_partial:
        mask = full_coverage_mask(src_mask, src_transform, params2, param_transform, kernel_shape)   # :495
        mask_us = gdal_restate.reproject_array(mask, param_transform, None, src_w.shape, src_transform, 0,
                                               'nearest')                   # :497
        _set_mask(param_us, mask_us.astype('bool', copy=False))              # :498
    else:
        _set_mask(param_us, src_mask)                                        # :500
    return apply_same_grid(src_w, param_us)                                  # :503


def srcspace_fit(src, src_transform, src_nodata, ref, ref_transform, ref_nodata, model, kernel_shape, find_r2=False,
                 r2_inpaint_thresh=0.25, mask_partial=False, downsampling='average', upsampling='cubic_spline'):
    """ SrcSpaceModel.fit, kernel_model.py:516-535; returns params on the source grid. """
    src_w, ref_w = as_working(src), as_working(ref)
    resampling = pick_resampling(res_of(ref_transform), res_of(src_transform), downsampling, upsampling)   # :518
    ref_us = gdal_restate.reproject_array(ref_w, ref_transform, ref_nodata, src_w.shape, src_transform, NODATA,
                                          resampling)                        # :520
    params = fit_same_grid(src_w, src_nodata, ref_us, NODATA, model, kernel_shape, find_r2, r2_inpaint_thresh)   # :524
    if mask_partial:
        mask = full_coverage_mask(valid_mask(ref_w, ref_nodata), ref_transform, params[:2], src_transform,
                                  kernel_shape)                              # :528-530
        _set_mask(params, mask.astype('bool', copy=False))                   # :531
    else:
        _set_mask(params, valid_mask(src_w, src_nodata))                     # :533
    return params


def srcspace_apply(src, params):
    """ SrcSpaceModel.apply == KernelModel.apply, kernel_model.py:442-463 """
    return apply_same_grid(src, params)


def fuse_band(src, src_transform, src_nodata, ref, ref_transform, ref_nodata, model, kernel_shape, proc_crs='ref',
              find_r2=False, r2_inpaint_thresh=0.25, mask_partial=False):
    """ One (band, block) of RasterFuse._process_block, fuse.py:304-307: fit then apply.  Returns (params, corr). """
    if proc_crs == 'src':
        params = srcspace_fit(src, src_transform, src_nodata, ref, ref_transform, ref_nodata, model, kernel_shape,
                              find_r2, r2_inpaint_thresh, mask_partial)
        return params, srcspace_apply(src, params)
    params = refspace_fit(src, src_transform, src_nodata, ref, ref_transform, ref_nodata, model, kernel_shape, find_r2,
                          r2_inpaint_thresh)
    corr = refspace_apply(src, src_transform, src_nodata, params, ref_transform, kernel_shape, mask_partial)
    return params, corr


# ---------------------------------------------------------------------------------------------------------------------
# block windows of RasterFuse / RasterPairReader for a single block per band
# ---------------------------------------------------------------------------------------------------------------------
def compare_band(src, src_transform, src_nodata, ref, ref_transform, ref_nodata, proc_crs='ref',
                 downsampling='average', upsampling='cubic_spline', dtype='float32') -> dict:
    """ get_block_sums, compare.py:232-254, for one band read as a single block (`block_windows`: the reader's
    boundless source window on whole reference pixels): re-project onto the processing grid (:236-241), then the
    masked sums. """
    src_w, src_transform, ref_w, ref_transform, _ = block_windows(src, src_transform, src_nodata, as_working(ref),
                                                                  ref_transform)      # :234 (self.read)
    if proc_crs == 'ref':
        resampling = pick_resampling(res_of(src_transform), res_of(ref_transform), downsampling, upsampling)   # :237
        src_w = gdal_restate.reproject_array(src_w, src_transform, src_nodata, ref_w.shape, ref_transform, NODATA,
                                             resampling)                     # :238
        src_nodata = NODATA
    else:
        resampling = pick_resampling(res_of(ref_transform), res_of(src_transform), downsampling, upsampling)   # :240
        ref_w = gdal_restate.reproject_array(ref_w, ref_transform, ref_nodata, src_w.shape, src_transform, NODATA,
                                             resampling)                     # :241
        ref_nodata = NODATA
    return compare_sums(src_w, src_nodata, ref_w, ref_nodata, dtype=dtype)


def _affine_mul(t, u):
    """ 6-coefficient affine product t * u. """
    ta, tb, tc, td, te, tf = [float(v) for v in t[:6]]
    ua, ub, uc, ud, ue, uf = [float(v) for v in u[:6]]
    return (ta * ua + tb * ud, ta * ub + tb * ue, ta * uc + tb * uf + tc,
            td * ua + te * ud, td * ub + te * ue, td * uc + te * uf + tf)


def _affine_inv(t):
    a, b, c, d, e, f = [float(v) for v in t[:6]]
    det = a * e - b * d
    ia, ib, id_, ie = e / det, -b / det, -d / det, a / det
    return (ia, ib, -(ia * c + ib * f), id_, ie, -(id_ * c + ie * f))


def _affine_pt(t, col, row):
    a, b, c, d, e, f = [float(v) for v in t[:6]]
    return a * col + b * row + c, d * col + e * row + f


def block_windows(src, src_transform, src_nodata, ref, ref_transform):
    """
    Single-block processing windows of RasterPairReader (raster_pair.py:292-296 with utils.expand_window_to_grid,
    utils.py:59-82) and the boundless block reads of RasterArray.from_rio_dataset (raster_array.py:175-199):

      ref window = whole reference pixels covering the source extent;
      src window = whole source pixels covering that reference window, read boundlessly (nodata outside the source).

    Returns (src_blk float32, src_blk_transform, ref_blk, ref_blk_transform, (row0, col0) of the source inside
    src_blk).
    """
    src_w = as_working(src)
    hs, ws = src_w.shape
    inv = _affine_inv(ref_transform)
    c0, r0 = _affine_pt(inv, *_affine_pt(src_transform, 0, 0))
    c1, r1 = _affine_pt(inv, *_affine_pt(src_transform, ws, hs))
    rc0, rr0, rc1, rr1 = int(np.floor(c0)), int(np.floor(r0)), int(np.ceil(c1)), int(np.ceil(r1))
    rc0, rr0 = max(rc0, 0), max(rr0, 0)
    rc1, rr1 = min(rc1, ref.shape[1]), min(rr1, ref.shape[0])
    ref_blk = np.array(ref[rr0:rr1, rc0:rc1], copy=True)
    ref_blk_tf = _affine_mul(ref_transform, (1, 0, rc0, 0, 1, rr0))
    sinv = _affine_inv(src_transform)
    sc0, sr0 = _affine_pt(sinv, *_affine_pt(ref_blk_tf, 0, 0))
    sc1, sr1 = _affine_pt(sinv, *_affine_pt(ref_blk_tf, ref_blk.shape[1], ref_blk.shape[0]))
    pc0, pr0 = min(int(np.floor(sc0 + 1e-9)), 0), min(int(np.floor(sr0 + 1e-9)), 0)
    pc1, pr1 = max(int(np.ceil(sc1 - 1e-9)), ws), max(int(np.ceil(sr1 - 1e-9)), hs)
    fill = NODATA if src_nodata is None else src_nodata
    src_blk = np.full((pr1 - pr0, pc1 - pc0), fill, dtype=src_w.dtype)
    src_blk[-pr0:-pr0 + hs, -pc0:-pc0 + ws] = src_w
    src_blk_tf = _affine_mul(src_transform, (1, 0, pc0, 0, 1, pr0))
    return src_blk, src_blk_tf, ref_blk, ref_blk_tf, (-pr0, -pc0)


def fuse_band_blocks(src, src_transform, src_nodata, ref, ref_transform, ref_nodata, model, kernel_shape,
                     proc_crs='ref', find_r2=False, r2_inpaint_thresh=0.25, mask_partial=False):
    """
    One band of RasterFuse.process with a single block (fuse.py:295-319): block windows -> fit -> apply -> crop the
    corrected block to the source extent (fuse.py:311).  Returns (params, param_transform, corr).
    """
    src_blk, src_blk_tf, ref_blk, ref_blk_tf, (r0, c0) = block_windows(src, src_transform, src_nodata, ref,
                                                                       ref_transform)
    hs, ws = np.shape(src)
    if proc_crs == 'src':
        params = srcspace_fit(src_blk, src_blk_tf, src_nodata, ref_blk, ref_blk_tf, ref_nodata, model, kernel_shape,
                              find_r2, r2_inpaint_thresh, mask_partial)
        corr = srcspace_apply(src_blk, params)
        return (np.ascontiguousarray(params[:, r0:r0 + hs, c0:c0 + ws]), tuple(src_transform[:6]),
                np.ascontiguousarray(corr[r0:r0 + hs, c0:c0 + ws]))
    params = refspace_fit(src_blk, src_blk_tf, src_nodata, ref_blk, ref_blk_tf, ref_nodata, model, kernel_shape,
                          find_r2, r2_inpaint_thresh)
    corr = refspace_apply(src_blk, src_blk_tf, src_nodata, params, ref_blk_tf, kernel_shape, mask_partial)
    return params, ref_blk_tf, np.ascontiguousarray(corr[r0:r0 + hs, c0:c0 + ws])

// upsample_poly.cuh -- interface of the fast cubic-spline up-sampling path (upsample_poly.cu) used by resample.cu.
#pragma once

#include "hb_common.cuh"

struct UpPolyGeom {
    long hs, ws;              // destination (fine) grid
    long hp, wp;              // coarse grid
    double sx, ox, sy, oy;    // destination pixel index -> coarse pixel coordinate: c = s * (i + 0.5) + o
    OutSpec ospec;            // output dtype / nodata of the corrected plane (apply mode); plain float32 otherwise
};

// the fast paths need >= ~1.6 destination pixels per coarse pixel and a destination width that is a multiple of 4
// (the caller also checks pointer / pitch alignment)
bool hb_up_poly_eligible(const UpPolyGeom &g);

// corr = up(gain) * src + up(offset)      (params: float32 [2][hp][wp], NaN = nodata; nd.ivalue = -1 if not an integer)
// `out` holds g.ospec.dtype elements (the conversion is fused into the stores)
int hb_up_poly_apply(const void *src, int src_dtype, NoData nd, const float *params, const UpPolyGeom &g, float *out,
                     cudaStream_t stream);
// plain up-sampling of ONE float32 band in double precision (out: [hs][ws]); two-band requests stay on the caller's
// general kernel
int hb_up_poly_resample(const float *coarse, int nb, const UpPolyGeom &g, float *out, cudaStream_t stream);

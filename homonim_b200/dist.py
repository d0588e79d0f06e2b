"""
Multi-GPU execution of the kernel-model path: one process per GPU, ``torch.distributed`` for the plumbing.

The reference parallelises over independent ``(band, block)`` units on a thread pool (fuse.py:396-408,
raster_pair.py:379-428); neighbouring blocks are only coupled through a fixed overlap of ``ceil(kernel / 2)`` proc-grid
pixels (utils.py:136-153).  Two regimes follow:

* **Batch / mosaic** (`shard_sources`): independent source images are dealt round-robin to the ranks; the reference
  image is replicated.  No data-path collective at all.
* **One raster as row bands** (`RowBands`, `fuse_refspace_sharded`, `fit_same_grid_sharded`): rank g owns a band of
  source rows whose edges coincide with proc-grid (reference) pixel rows.

  - proc_crs = ref: every rank down-samples its own source rows, the proc-grid planes (3000 x 3000 float32 = 36 MB for
    the 60k x 60k configuration -- 1/400 of the source) are **all-gathered** so that the block normalisation of
    gain-blk-offset (a whole-block statistic, kernel_model.py:216-229) stays global; every rank then fits its own proc
    rows plus a halo (`halo_rows`) and applies the parameters to its own source rows.  Results equal the single-GPU
    results up to the summation order of the fit kernel's running sums (> 99.9 % of the parameters bit-identical).
  - same grid (proc_crs = src): the window sums need ``kh // 2`` rows of source and reference from each neighbour:
    `exchange_halos` does that with point-to-point send / recv between row-band neighbours (NCCL P2P over NVLink on
    GPUs, gloo in the CPU tests); rows beyond the raster stay absent, which is the reference's zero padding
    (cv BORDER_CONSTANT).  The fit then runs on the extended band and the halo rows of the result are dropped.

All timing of multi-GPU runs is done by the caller on the device (bench.py: CUDA events, max over ranks).
"""
from typing import List, NamedTuple, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from homonim_b200.geometry import Affine
from homonim_b200.raster_array import RasterArray


def shard_sources(n_sources: int, rank: int, world_size: int) -> List[int]:
    """ Indexes of the independent source images this rank processes (round-robin; SURVEY.md 8e batch mode). """
    return list(range(rank, n_sources, world_size))


class RowBands(NamedTuple):
    """
    Partition of ``n_rows`` proc-grid rows into ``world_size`` contiguous bands, as even as possible.
    ``starts[g] .. starts[g + 1]`` is rank g's band (empty bands are allowed when there are more ranks than rows).
    """
    starts: Tuple[int, ...]

    @classmethod
    def split(cls, n_rows: int, world_size: int) -> 'RowBands':
        base, extra = divmod(int(n_rows), int(world_size))
        starts = [0]
        for g in range(world_size):
            starts.append(starts[-1] + base + (1 if g < extra else 0))
        return cls(tuple(starts))

    def band(self, rank: int) -> Tuple[int, int]:
        return self.starts[rank], self.starts[rank + 1]

    def size(self, rank: int) -> int:
        return self.starts[rank + 1] - self.starts[rank]

    def with_halo(self, rank: int, halo: int) -> Tuple[int, int]:
        """ Rank's band extended by ``halo`` rows on both sides, clipped to the raster. """
        a, b = self.band(rank)
        return max(a - halo, 0), min(b + halo, self.starts[-1])


def halo_rows(kernel_shape: Sequence[int], proc_crs_ref: bool, inpaint: bool) -> int:
    """
    Rows of proc-grid halo a row band needs on each side for its results to equal the whole-raster results:
    ``kh // 2`` for the window sums (the reference's block overlap uses the safe ``ceil(kh / 2)``, utils.py:136-153),
    + 2 for the cubic-spline taps of the first / last up-sampled rows (proc_crs = ref), + 100 when in-painting is on
    (fillnodata's search radius, kernel_model.py:366).
    """
    return int(kernel_shape[0]) // 2 + (2 if proc_crs_ref else 0) + (100 if inpaint else 0)


def exchange_halos(local: torch.Tensor, bands: RowBands, halo: int, group=None) -> Tuple[torch.Tensor, int]:
    """
    Point-to-point halo exchange between row-band neighbours.

    ``local``: this rank's rows ``[..., n_local, width]`` of a raster split by ``bands``.  Returns ``(extended, top)``
    where ``extended`` holds the rows ``bands.with_halo(rank, halo)`` and ``top`` is the number of halo rows that were
    prepended.  A halo deeper than a neighbour's band is served by the next ranks as well.
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a, b = bands.band(rank)
    lo, hi = bands.with_halo(rank, halo)
    if world == 1 or halo == 0:
        return local, 0
    ops, recv_bufs = [], []
    lead = local.shape[:-2]
    width = local.shape[-1]
    for peer in range(world):
        if peer == rank:
            continue
        pa, pb = bands.band(peer)
        plo, phi = bands.with_halo(peer, halo)
        # rows of mine that the peer needs
        s0, s1 = max(a, plo), min(b, phi)
        if s1 > s0:
            chunk = local[..., s0 - a:s1 - a, :].contiguous()
            ops.append(dist.P2POp(dist.isend, chunk, peer, group))
        # rows of the peer that I need
        r0, r1 = max(pa, lo), min(pb, hi)
        if r1 > r0:
            buf = torch.empty(lead + (r1 - r0, width), dtype=local.dtype, device=local.device)
            recv_bufs.append((r0, buf))
            ops.append(dist.P2POp(dist.irecv, buf, peer, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    pieces = sorted(recv_bufs + [(a, local)], key=lambda t: t[0])
    extended = torch.cat([p for _, p in pieces], dim=-2) if len(pieces) > 1 else local
    return extended, a - lo


def all_gather_rows(local: torch.Tensor, bands: RowBands, group=None) -> torch.Tensor:
    """ All-gather row bands of unequal height into the whole ``[..., n_rows, width]`` plane on every rank. """
    world = dist.get_world_size(group)
    if world == 1:
        return local
    lead, width = local.shape[:-2], local.shape[-1]
    max_rows = max(bands.size(g) for g in range(world))
    padded = torch.zeros(lead + (max_rows, width), dtype=local.dtype, device=local.device)
    padded[..., :local.shape[-2], :] = local
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    return torch.cat([gathered[g][..., :bands.size(g), :] for g in range(world)], dim=-2)


def source_band_for_proc_rows(src_ra_shape: Tuple[int, int], src_transform: Affine, ref_transform: Affine,
                              proc_rows: Tuple[int, int]) -> Tuple[int, int]:
    """
    Source rows whose pixel CENTRES fall inside proc-grid rows ``[proc_rows[0], proc_rows[1])`` (north-up grids).  With
    an integer, aligned ratio these are exactly the ``ratio * n`` source rows under the band.
    """
    import math
    sy = ref_transform.e / src_transform.e
    oy = (ref_transform.f - src_transform.f) / src_transform.e
    r0 = max(int(math.ceil(sy * proc_rows[0] + oy - 0.5)), 0)
    r1 = min(int(math.ceil(sy * proc_rows[1] + oy - 0.5)), src_ra_shape[0])
    return r0, max(r1, r0)


def fit_row_window(bands: RowBands, rank: int, halo: int) -> Tuple[int, int, int, int]:
    """
    Proc-grid rows of the sharded proc_crs = ref path for ``rank``: ``(lo, hi, plo, phi)``.  ``[lo, hi)`` are the rows the
    rank fits (its band extended by ``halo`` rows, clipped to the raster); ``[plo, phi)`` are the rows of that fit it keeps
    and hands to the up-sampler (its band plus the 2 rows of cubic-spline support on either side, clipped).  Every kept
    row is at least ``halo - 2`` rows away from a cut edge of the fitted window -- the distance over which window sums
    (and the in-painting search) are affected by the cut -- unless that edge is the raster's own.
    """
    a, b = bands.band(rank)
    hp = bands.starts[-1]
    lo, hi = max(a - halo, 0), min(b + halo, hp)
    plo, phi = max(a - 2, 0), min(b + 2, hp)
    return lo, hi, plo, phi


def fuse_refspace_sharded(model, src_local: RasterArray, ref_ra: RasterArray, bands: RowBands, group=None, out=None
                          ) -> Tuple[RasterArray, RasterArray]:
    """
    proc_crs = ref fit + apply of ONE band of a raster that is sharded by rows (configuration C5a).

    ``src_local`` holds this rank's source rows (its transform already points at its first row); ``ref_ra`` is the
    whole (replicated) reference band on the proc grid; ``bands`` partitions the proc-grid rows.  Returns
    ``(corr_local, param_ra)``: the corrected rows of this rank and the parameters of the proc-grid rows they depend on
    (this rank's band plus the 2 rows of cubic-spline support on either side; ``param_ra.transform`` points at them).
    ``out`` (optional): float32 CUDA tensor to receive the corrected rows.

    Only the down-sampled proc-grid plane -- 1/ratio^2 of the source -- crosses the interconnect: it is all-gathered so
    that the block normalisation of gain-blk-offset (a whole-block statistic, kernel_model.py:216-229) stays global;
    every rank then fits just its own proc rows plus the halo that makes them equal the whole-raster fit
    (`halo_rows`: window half-height + spline support + the in-painting search radius).
    """
    from homonim_b200 import kernel_model as km
    from homonim_b200.enums import Model
    rank = dist.get_rank(group)
    a, b = bands.band(rank)
    nan = float('nan')
    src_t = km._to_device(src_local.array)
    # 1. down-sample my source rows onto my proc rows
    local_tf = ref_ra.transform * Affine.translation(0, a)
    src_ds_local = km._downsample_average(src_t, src_local.transform, src_local.nodata, (b - a, ref_ra.width), local_tf)
    # 2. the proc-grid plane is tiny: gather it everywhere
    src_ds = all_gather_rows(src_ds_local, bands, group)
    ref_t = km._to_device(ref_ra.array)
    # 3. whole-block statistics (redundant on every rank: three passes over the small proc-grid planes)
    norm = None
    if model.model == Model.gain_blk_offset:
        norm = model._block_norm(src_ds, nan, ref_t, ref_ra.nodata)
    # 4. fit my proc rows + halo; keep the rows my source rows' spline taps touch
    inpaint = model.model == Model.gain_offset and model._r2_inpaint_thresh is not None
    halo = halo_rows(model.kernel_shape, proc_crs_ref=True, inpaint=inpaint)
    lo, hi, plo, phi = fit_row_window(bands, rank, halo)
    params_ext = model._fit_planes(src_ds[lo:hi], nan, ref_t[lo:hi], ref_ra.nodata, norm=norm)
    params = params_ext[:, plo - lo:phi - lo].contiguous()
    param_ra = RasterArray(params, ref_ra.crs, ref_ra.transform * Affine.translation(0, plo), nodata=nan)
    # 5. apply to my source rows
    corr_local = model.apply(RasterArray(src_t, src_local.crs, src_local.transform, nodata=src_local.nodata), param_ra,
                             out=out)
    return corr_local, param_ra


def fit_same_grid_sharded(model, src_local: torch.Tensor, src_nodata, ref_local: torch.Tensor, ref_nodata,
                          bands: RowBands, group=None) -> torch.Tensor:
    """
    Same-grid fit of a raster sharded by rows (configuration C5b): halo exchange of ``kh // 2`` rows of both planes
    with the row-band neighbours, fit on the extended band, halo rows dropped.  gain and gain-offset models only --
    gain-blk-offset needs the block statistics of the whole raster (use `fuse_refspace_sharded`, or fit unsharded).
    """
    from homonim_b200.enums import Model
    if model.model == Model.gain_blk_offset:
        raise NotImplementedError('sharded same-grid fitting of gain-blk-offset needs whole-raster block statistics')
    halo = halo_rows(model.kernel_shape, proc_crs_ref=False,
                     inpaint=(model.model == Model.gain_offset and model._r2_inpaint_thresh is not None))
    rank = dist.get_rank(group)
    src_ext, top = exchange_halos(src_local, bands, halo, group)
    ref_ext, _ = exchange_halos(ref_local, bands, halo, group)
    params = model._fit_planes(src_ext.contiguous(), src_nodata, ref_ext.contiguous(), ref_nodata)
    return params[:, top:top + bands.size(rank), :].contiguous()

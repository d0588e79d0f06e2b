"""
CPU tier: the multi-GPU host logic of homonim_b200/dist.py with world_size 2 on the gloo backend -- row-band
partitioning, halo sizes, halo exchange between row-band neighbours and the proc-grid all-gather.  (The kernels
themselves need a GPU; `tests/test_dist_gpu.py` covers sharded == unsharded there.)
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from homonim_b200.dist import RowBands, all_gather_rows, alloc_with_halo, block_norm_sharded, exchange_halos, \
    exchange_halos_inplace, halo_rows, shard_sources, source_band_for_proc_rows
from homonim_b200 import Affine


def test_row_bands_and_shards():
    b = RowBands.split(10, 4)
    assert b.starts == (0, 3, 6, 8, 10) and [b.size(g) for g in range(4)] == [3, 3, 2, 2]
    assert b.with_halo(0, 2) == (0, 5) and b.with_halo(2, 9) == (0, 10) and b.with_halo(3, 1) == (7, 10)
    assert RowBands.split(3, 8).starts[-1] == 3 and sum(RowBands.split(3, 8).size(g) for g in range(8)) == 3
    assert shard_sources(64, 3, 8) == list(range(3, 64, 8)) and shard_sources(2, 5, 8) == []
    assert sorted(sum((shard_sources(13, r, 4) for r in range(4)), [])) == list(range(13))


def test_halo_rows_follow_the_reference_overlap_rule():
    # kh // 2 for the window sums, + 2 spline taps (proc_crs=ref), + 100 for fillnodata's search radius
    assert halo_rows((15, 15), proc_crs_ref=True, inpaint=False) == 9
    assert halo_rows((15, 15), proc_crs_ref=False, inpaint=False) == 7
    assert halo_rows((5, 31), proc_crs_ref=True, inpaint=True) == 104
    assert halo_rows((1, 1), proc_crs_ref=False, inpaint=False) == 0


def test_source_band_for_proc_rows():
    ref_tf = Affine(10, 0, 0, 0, -10, 0)
    src_tf = Affine(0.5, 0, 0, 0, -0.5, 0)
    assert source_band_for_proc_rows((10000, 10000), src_tf, ref_tf, (0, 125)) == (0, 2500)
    assert source_band_for_proc_rows((10000, 10000), src_tf, ref_tf, (125, 500)) == (2500, 10000)
    shifted = Affine(0.5, 0, 0, 0, -0.5, -3.0)          # source starts 6 source rows below the proc grid's top
    assert source_band_for_proc_rows((100, 10), shifted, ref_tf, (0, 1)) == (0, 14)
    assert source_band_for_proc_rows((100, 10), shifted, ref_tf, (1, 2)) == (14, 34)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, halo, result_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        full = torch.arange(2 * n_rows * 7, dtype=torch.float32).reshape(2, n_rows, 7)
        bands = RowBands.split(n_rows, world)
        a, b = bands.band(rank)
        local = full[:, a:b, :].contiguous()
        ext, top = exchange_halos(local, bands, halo)
        lo, hi = bands.with_halo(rank, halo)
        assert top == a - lo
        assert torch.equal(ext, full[:, lo:hi, :]), f'rank {rank}: halo exchange mismatch'
        gathered = all_gather_rows(local[0].contiguous(), bands)
        assert torch.equal(gathered, full[0]), f'rank {rank}: all-gather mismatch'
        # the in-place variant: own rows written into a plane allocated with its halo, the rest received in place
        buf, top2 = alloc_with_halo(bands, rank, halo, 7, torch.float32, 'cpu')
        assert top2 == top and buf.shape[0] == hi - lo
        buf.fill_(-1.0)
        buf[top2:top2 + (b - a)] = full[0, a:b]
        exchange_halos_inplace(buf, bands, halo)
        assert torch.equal(buf, full[0, lo:hi]), f'rank {rank}: in-place halo exchange mismatch'
        # max-over-ranks reduction used for timing
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        open(os.path.join(result_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_rows, halo', [(11, 2), (6, 4), (3, 1)])
def test_halo_exchange_and_gather_world_size_2(tmp_path, n_rows, halo):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_rows, halo, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))


def test_halo_exchange_world_size_3_deep_halo(tmp_path):
    # halo deeper than a neighbour's band: rows must also come from the next rank
    world = 3
    mp.spawn(_worker, args=(world, _free_port(), 7, 4, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))


def test_fit_row_window_margins():
    """ Row windows of the sharded proc_crs = ref path: the rows a rank keeps lie inside the rows it fits, far enough
    from every cut edge that the cut cannot influence them, and together the kept bands cover every rank's taps. """
    import random
    from homonim_b200.dist import RowBands, fit_row_window, halo_rows
    rnd = random.Random(3)
    for _ in range(300):
        hp = rnd.randint(1, 400)
        world = rnd.randint(1, 9)
        kh = rnd.choice([1, 3, 5, 15, 31])
        inpaint = rnd.random() < 0.3
        halo = halo_rows((kh, kh), proc_crs_ref=True, inpaint=inpaint)
        bands = RowBands.split(hp, world)
        for rank in range(world):
            a, b = bands.band(rank)
            lo, hi, plo, phi = fit_row_window(bands, rank, halo)
            assert 0 <= lo <= plo <= phi <= hi <= hp
            if b > a:
                # the up-sampler's taps for source rows under proc rows [a, b) are proc rows [a - 2, b + 2)
                assert plo == max(a - 2, 0) and phi == min(b + 2, hp)
                # distance from the kept rows to a CUT edge (an edge that is not the raster's own)
                if lo > 0:
                    assert plo - lo >= halo - 2
                if hi < hp:
                    assert hi - phi >= halo - 2


# ---------------------------------------------------------------------------------------------------------------------
# whole-block statistics of a sharded block: the accumulate / all-gather / merge protocol of block_norm_sharded with a
# numpy stand-in for the two native steps (same 12 + 12 + 8 bit radix select, same accumulators)
# ---------------------------------------------------------------------------------------------------------------------
class _NumpyBlockNorm:
    """ Stand-in for dist.NativeBlockNorm: what hb_block_norm_partial / hb_block_norm_merge compute, in numpy. """
    BINS = 4096

    def __init__(self, src_local, src_nodata, ref_local, ref_nodata, n_local_max, n_total, rank):
        self.sizes = (n_local_max, n_total, rank)
        s, r = src_local.numpy().ravel(), ref_local.numpy().ravel()
        valid = ~np.isnan(s) & ~np.isnan(r)
        self.s, self.r = s[valid], r[valid]
        _NumpyBlockNorm.last_sizes = self.sizes
        self.norm = torch.zeros(2, dtype=torch.float64)
        self.n, self.mean, self.prefix, self.rank, self.gamma = 0, [0.0, 0.0], [0] * 4, [0] * 4, 0.0

    @staticmethod
    def _keys(v):
        b = v.view(np.uint32).astype(np.uint64)
        return np.where(b & 0x80000000, (~b) & 0xffffffff, b | 0x80000000).astype(np.uint64)

    def partial(self, level):
        acc = np.zeros(5 + 4 * self.BINS, dtype=np.float64)          # n, sum[2], ssd[2], hist[4][BINS] (exact in f64)
        ks, kr = self._keys(self.s), self._keys(self.r)
        hist = acc[5:].reshape(4, self.BINS)
        if level == 0:
            acc[0], acc[1], acc[2] = self.s.size, self.s.sum(dtype=np.float64), self.r.sum(dtype=np.float64)
            hist[0] = np.bincount((ks >> 20).astype(np.int64), minlength=self.BINS)
            hist[2] = np.bincount((kr >> 20).astype(np.int64), minlength=self.BINS)
        else:
            if level == 1:
                acc[3] = ((self.s.astype(np.float64) - self.mean[0]) ** 2).sum()
                acc[4] = ((self.r.astype(np.float64) - self.mean[1]) ** 2).sum()
            sh, bits = (20, 12) if level == 1 else (8, 8)
            for q in range(4):
                k = ks if q < 2 else kr
                sel = k[(k >> sh) == self.prefix[q]]
                bins = ((sel >> 8) & 0xfff) if level == 1 else (sel & 0xff)
                hist[q, :1 << bits] = np.bincount(bins.astype(np.int64), minlength=1 << bits)
        return torch.from_numpy(acc.view(np.uint8).copy())

    def merge(self, level, gathered, world):
        acc = gathered.numpy().view(np.float64).reshape(world, -1).sum(axis=0)
        hist = acc[5:].reshape(4, self.BINS)
        if level == 0:
            self.n = int(acc[0])
            if self.n == 0:
                return
            self.mean = [acc[1] / self.n, acc[2] / self.n]
            q32 = np.float32(1) / np.float32(100)
            vi = np.float32(np.float32(self.n - 1) * q32)          # numpy's 'linear' method: (n - 1) * q in float32
            k0 = int(np.floor(vi))
            self.gamma = np.float32(float(vi) - k0)
            last = self.n - 1
            k1 = k0 + 1
            if vi >= np.float32(last):
                k0 = k1 = last
            self.rank = [max(k0, 0), min(max(k1, 0), max(last, 0))] * 2
        elif level == 1:
            self.ssd = [acc[3], acc[4]]
        if self.n == 0:
            return
        for q in range(4):
            h = hist[0 if (level == 0 and q < 2) else (2 if level == 0 else q)]
            c = np.cumsum(h)
            b = int(np.searchsorted(c, self.rank[q], side='right'))
            self.rank[q] -= int(c[b - 1]) if b > 0 else 0
            self.prefix[q] = b if level == 0 else ((self.prefix[q] << (12 if level == 1 else 8)) | b)
        if level == 2:
            def val(key):
                b = np.uint32(key & 0x7fffffff) if key & 0x80000000 else np.uint32(~np.uint32(key))
                return np.array([b], dtype=np.uint32).view(np.float32)[0]
            std_s = np.float32(np.sqrt(self.ssd[0] / self.n))
            std_r = np.float32(np.sqrt(self.ssd[1] / self.n))
            n0 = float(std_r / std_s)
            p = []
            for a in range(2):
                lo, hi = val(self.prefix[2 * a]), val(self.prefix[2 * a + 1])
                d = np.float32(hi - lo)
                v = np.float32(lo + np.float32(d * self.gamma))
                if self.gamma >= 0.5:
                    v = np.float32(hi - np.float32(d * np.float32(np.float32(1) - self.gamma)))
                p.append(v)
            self.norm[0], self.norm[1] = n0, float(p[1]) - float(p[0]) * n0


def _norm_worker(rank, world, port, result_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        h, w = 97, 61
        src = (rng.normal(size=(h, w)) * 0.1 + 0.3).astype('float32')
        ref = (0.7 * src + 0.05 + rng.normal(size=(h, w)) * 0.01).astype('float32')
        src[rng.random((h, w)) < 0.05] = np.nan
        ref[5:9, 3:30] = np.nan
        bands = RowBands.split(h, world)
        a, b = bands.band(rank)
        norm = block_norm_sharded(torch.from_numpy(src[a:b].copy()), float('nan'), torch.from_numpy(ref[a:b].copy()),
                                  float('nan'), backend=_NumpyBlockNorm)
        # (the shard sizes the native back end needs were found with one small all-reduce)
        assert _NumpyBlockNorm.last_sizes == (max(bands.size(g) for g in range(world)) * w, h * w, rank)
        mask = ~np.isnan(src) & ~np.isnan(ref)
        exp0 = np.std(ref[mask]) / np.std(src[mask])
        exp1 = np.percentile(ref[mask], 1) - np.percentile(src[mask], 1) * exp0
        # order statistics are exact, so the percentile part agrees to the last bit given the same n0; the float32
        # np.std of numpy is pairwise-summed, the protocol's is double-accumulated: within 1 float32 ulp
        assert abs(float(norm[0]) - float(exp0)) <= 2e-7 * abs(float(exp0)), (norm, exp0)
        assert abs(float(norm[1]) - float(exp1)) <= 1e-6 * max(abs(float(exp1)), 1e-3), (norm, exp1)
        # every rank must hold the SAME statistics
        both = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(both, norm)
        assert all(torch.equal(both[0], t) for t in both)
        open(os.path.join(result_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_block_norm_sharded_protocol(tmp_path, world):
    mp.spawn(_norm_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))

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

from homonim_b200.dist import RowBands, all_gather_rows, exchange_halos, halo_rows, shard_sources, \
    source_band_for_proc_rows
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

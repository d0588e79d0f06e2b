"""
GPU tier: row-band sharding (homonim_b200/dist.py) must reproduce the single-GPU result -- always with MORE THAN ONE
rank.  With 2+ GPUs every rank drives its own GPU over NCCL (P2P halos, all-gathered block statistics).  On a
single-GPU box the same two-rank job runs with both ranks on cuda:0 over gloo (NCCL refuses two ranks on one device;
dist.py then stages the few exchanged rows / accumulators through host memory): every kernel of the sharded path --
partial / merged block statistics, row-range fits inside a halo, the shifted up-sampler -- still runs on the GPU and is
compared with the unsharded result.
"""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip('torch')
import torch.distributed as dist   # noqa: E402
import torch.multiprocessing as mp   # noqa: E402

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_dir, shared_gpu):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dev = 0 if shared_gpu else rank
    torch.cuda.set_device(dev)
    if shared_gpu:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    else:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', dev))
    rank_dev = dev
    try:
        from homonim_b200 import Affine, KernelModel, Model, RasterArray, RefSpaceModel
        from homonim_b200.dist import RowBands, alloc_with_halo, block_norm_sharded, fit_apply_same_grid_sharded, \
            fit_same_grid_sharded, fuse_refspace_sharded, halo_rows, source_band_for_proc_rows
        from homonim_b200.synthetic import make_pair
        nan = float('nan')
        # ---- proc_crs = ref, one raster as row bands (C5a geometry, scaled down) -----------------------------------
        src_ra, ref_ra = make_pair(90, 64, 20, bands=1, dtype='float32', mu=0.3, seed=21, device=f'cuda:{rank_dev}',
                                   src_nodata=nan, ref_pad=0)
        src = RasterArray(src_ra.array[0].contiguous(), src_ra.crs, src_ra.transform, nodata=nan)
        ref = RasterArray(ref_ra.array[0].contiguous(), ref_ra.crs, ref_ra.transform, nodata=nan)
        for model_name, kshape, thresh in ((Model.gain_blk_offset, (15, 15), None), (Model.gain_offset, (5, 5), 0.25)):
            model = RefSpaceModel(model_name, kshape, find_r2=True, r2_inpaint_thresh=thresh)
            full_params = model.fit(src, ref)
            full_corr = model.apply(src, full_params).array
            bands = RowBands.split(ref.height, world)
            r0, r1 = source_band_for_proc_rows(src.shape, src.transform, ref.transform, bands.band(rank))
            src_local = RasterArray(src.array[r0:r1].contiguous(), src.crs, src.transform * Affine.translation(0, r0),
                                    nodata=nan)
            corr_local, param_ra = fuse_refspace_sharded(model, src_local, ref, bands)
            # the rank's parameters: its proc rows plus 2 rows of spline support on either side
            a, b = bands.band(rank)
            plo, phi = max(a - 2, 0), min(b + 2, ref.height)
            assert param_ra.transform == ref.transform * Affine.translation(0, plo)
            exp_p, got_p = full_params.array[:, plo:phi], param_ra.array
            assert torch.equal(torch.isnan(got_p), torch.isnan(exp_p))
            same = ((got_p == exp_p) | (torch.isnan(got_p) & torch.isnan(exp_p))).float().mean().item()
            assert same > 0.999, f'sharded params: only {same:.5f} bit-identical'
            exp_c, got_c = full_corr[r0:r1], corr_local.array
            assert torch.equal(torch.isnan(got_c), torch.isnan(exp_c)), 'sharded corr masks differ'
            fin = torch.isfinite(exp_c)
            rel = ((got_c - exp_c).abs()[fin] / exp_c.abs()[fin].clamp_min(1e-3 * exp_c[fin].abs().mean())).max().item()
            assert rel <= 1e-4, rel
        # ---- same grid (C5b geometry, scaled down): halo exchange of kh // 2 rows ------------------------------------
        s_ra, r_ra = make_pair(400, 333, 1, bands=1, dtype='float32', mu=0.3, seed=22, device=f'cuda:{rank_dev}',
                               src_nodata=nan, ref_pad=0)
        s_full, r_full = s_ra.array[0].contiguous(), r_ra.array[0].contiguous()
        km = KernelModel(Model.gain_offset, (15, 15), find_r2=True, r2_inpaint_thresh=None)
        full = km._fit_planes(s_full, nan, r_full, nan)
        bands = RowBands.split(s_full.shape[0], world)
        a, b = bands.band(rank)
        part = fit_same_grid_sharded(km, s_full[a:b].contiguous(), nan, r_full[a:b].contiguous(), nan, bands)
        ref_part = full[:, a:b]
        assert torch.equal(torch.isnan(part), torch.isnan(ref_part))
        same = ((part == ref_part) | (torch.isnan(part) & torch.isnan(ref_part))).float().mean().item()
        assert same > 0.999, f'same-grid sharded fit: only {same:.5f} bit-identical'
        fin = torch.isfinite(ref_part)
        rel = ((part - ref_part).abs()[fin] / ref_part.abs()[fin].clamp_min(1e-3)).max().item()
        assert rel <= 1e-4, rel
        # ---- whole-block statistics from per-rank accumulators == hb_block_norm of the whole planes --------------------
        norm_full = km._block_norm(s_full, nan, r_full, nan)
        norm_sh = block_norm_sharded(s_full[a:b], nan, r_full[a:b], nan)
        nf, ns = norm_full.cpu().numpy(), norm_sh.cpu().numpy()
        assert np.array_equal(nf, ns), (nf, ns)          # (numpy's pairwise np.std replayed across the shards: identical)
        # ---- C5b: same grid, gain-blk-offset 15 x 15 (BASELINE configs[4]): statistics merged over the ranks, halo rows
        #      received in place, fit + apply in one kernel on the rank's rows ----------------------------------------------
        for model_name in (Model.gain_blk_offset, Model.gain_offset):
            kb = KernelModel(model_name, (15, 15), r2_inpaint_thresh=None)
            full_p = kb._fit_planes(s_full, nan, r_full, nan)
            full_c = kb._apply_planes(s_full, nan, full_p, mask_src=False)
            halo = halo_rows((15, 15), proc_crs_ref=False, inpaint=False)
            s_ext, top = alloc_with_halo(bands, rank, halo, s_full.shape[1], torch.float32, s_full.device)
            r_ext, _ = alloc_with_halo(bands, rank, halo, s_full.shape[1], torch.float32, s_full.device)
            s_ext.fill_(-7.0); r_ext.fill_(-7.0)                # (halo rows must come from the neighbours)
            s_ext[top:top + (b - a)] = s_full[a:b]
            r_ext[top:top + (b - a)] = r_full[a:b]
            corr = fit_apply_same_grid_sharded(kb, s_ext, nan, r_ext, nan, bands)
            exp = full_c[a:b]
            assert torch.equal(torch.isnan(corr), torch.isnan(exp)), f'{model_name}: sharded corr masks differ'
            fin = torch.isfinite(exp)
            rel = ((corr - exp).abs()[fin] / exp.abs()[fin].clamp_min(1e-3 * exp[fin].abs().mean())).max().item()
            assert rel <= 1e-4, (model_name, rel)
            same = ((corr == exp) | (torch.isnan(corr) & torch.isnan(exp))).float().mean().item()
            assert same > 0.99, f'{model_name}: only {same:.5f} of the corrected pixels bit-identical'
            # and the parameter form of the same thing
            part = fit_same_grid_sharded(kb, s_full[a:b].contiguous(), nan, r_full[a:b].contiguous(), nan, bands)
            assert torch.equal(torch.isnan(part), torch.isnan(full_p[:, a:b]))
        open(os.path.join(result_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


def test_sharded_equals_unsharded(tmp_path):
    n_gpu = torch.cuda.device_count()
    assert n_gpu >= 1
    shared_gpu = n_gpu < 2
    world = 2 if shared_gpu else min(n_gpu, 4)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), shared_gpu), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))

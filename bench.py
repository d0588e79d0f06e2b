#!/usr/bin/env python
"""
bench.py -- fit + apply throughput of the kernel-model hot path (BASELINE.json metric) on N B200 GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c2-gain|c3|c4|c5a]

One "step" is one pass of the hot path over one synthetic source image: RasterFuse.process() = per band
RefSpaceModel/SrcSpaceModel .fit() + .apply().  Default workload (N = 1): BASELINE.json configs[1] --
a 4-band uint16 10 000 x 10 000 aerial image against a 10 m reference (20x coarser), gain-offset 15x15 with
r2_inpaint_thresh = 0.25, proc_crs = ref.  With N > 1 every rank corrects its own source image (the batch-mosaic
regime, no data-path collective): weak scaling, value = pixels of all ranks / max-over-ranks time.  `--workload c5a` is
the one-raster regime instead: ONE 60 000 x 60 000 4-band float32 raster split into row bands over the ranks (strong
scaling; the down-sampled proc-grid planes are all-gathered over NCCL, homonim_b200/dist.py).

The JSON line carries: `value` (device-resident inputs, CUDA-event timed), `e2e` (the same call with HOST buffers:
pinned host -> device copies of the inputs and device -> host copy of the corrected image inside the timed region),
`roofline` for the dominant kernel (per-launch CUDA-event durations collected live inside the timed region),
`cpu_baseline` (the oracle port -- the reference's own cv2 + numpy algorithm -- on a bounded sample, host cores), and
`clocks` sampled with nvidia-smi during the timed region.

`--impl reference` times the reference's CPU implementation of the same path (the oracle port: /root/reference does
not exist on the GPU box and rasterio/GDAL are not installable) on the host cores, on the same workload definition.
"""
import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

REPO = pathlib.Path(__file__).resolve().parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

import numpy as np  # noqa: E402

NAN = float('nan')

WORKLOADS = {
    # name: hp, wp (proc grid), ratio, bands, dtype, mu, src_nodata, model, kernel, thresh, proc_crs
    'c2': dict(hp=500, wp=500, ratio=20, bands=4, dtype='uint16', mu=3000.0, src_nodata=0.0, model='gain-offset',
               kernel_shape=(15, 15), r2_inpaint_thresh=0.25, proc_crs='ref',
               desc='C2: synthetic 4-band uint16 10000x10000 aerial vs 10 m reference (20x coarser), gain-offset '
                    '15x15, r2_inpaint_thresh=0.25, proc_crs=ref'),
    'c2-gain': dict(hp=500, wp=500, ratio=20, bands=4, dtype='uint16', mu=3000.0, src_nodata=0.0, model='gain',
                    kernel_shape=(1, 1), r2_inpaint_thresh=0.25, proc_crs='ref',
                    desc='C2: synthetic 4-band uint16 10000x10000 aerial vs 10 m reference, gain 1x1, proc_crs=ref'),
    'c3': dict(hp=10000, wp=10000, ratio=2, bands=4, dtype='float32', mu=0.3, src_nodata=NAN, model='gain-offset',
               kernel_shape=(31, 31), r2_inpaint_thresh=None, proc_crs='src',
               desc='C3: synthetic 4-band float32 20000x20000, proc_crs=src (fit at source resolution), gain-offset '
                    '31x31, no in-painting'),
    'tiny': dict(hp=40, wp=36, ratio=8, bands=2, dtype='uint16', mu=3000.0, src_nodata=0.0, model='gain-blk-offset',
                 kernel_shape=(5, 5), r2_inpaint_thresh=0.25, proc_crs='ref',
                 desc='tiny: synthetic 2-band uint16 320x288 (contract / smoke tests of this script, not a benchmark)'),
    'c4': dict(hp=400, wp=400, ratio=20, bands=4, dtype='uint16', mu=3000.0, src_nodata=0.0, model='gain-blk-offset',
               kernel_shape=(5, 5), r2_inpaint_thresh=0.25, proc_crs='ref',
               desc='C4: batch mosaic, synthetic 4-band uint16 8000x8000 sources vs a 10 m reference, one source per '
                    'GPU and step, gain-blk-offset 5x5, proc_crs=ref'),
    'c5a': dict(hp=3000, wp=3000, ratio=20, bands=4, dtype='float32', mu=0.3, src_nodata=NAN,
                model='gain-blk-offset', kernel_shape=(15, 15), r2_inpaint_thresh=0.25, proc_crs='ref', sharded=True,
                desc='C5a: ONE synthetic 4-band float32 60000x60000 raster sharded as row bands over the GPUs, '
                     'gain-blk-offset 15x15, proc_crs=ref (proc-grid planes all-gathered over NCCL)'),
}


def _clock_sampler(stop, samples, gpu_index, uuid=None):
    """
    SM clock / throttle-reason samples of one GPU while `stop` is unset.  NVML in-process (a query costs ~50 us, so the
    few-millisecond timed region gets several samples); falls back to polling nvidia-smi (~100 ms per query).
    Every sample: [sm_mhz, sm_max_mhz, power_w, hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap,
    perf_counter timestamp].
    """
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = None
        if uuid:
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if str(uuid).startswith('GPU-') else f'GPU-{uuid}')
            except Exception:
                handle = None
        if handle is None:
            handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        smax = pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)
        get_reasons = getattr(pynvml, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = [0x8, 0x40, 0x20, 0x4]        # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
        while not stop.is_set():
            sm = pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
            try:
                power = pynvml.nvmlDeviceGetPowerUsage(handle) / 1000.0
            except Exception:
                power = 0.0
            reasons = int(get_reasons(handle))
            samples.append([str(sm), str(smax), f'{power:.1f}'] +
                           ['Active' if reasons & b else 'Not Active' for b in bits] + [time.perf_counter()])
            stop.wait(0.002)
        return
    except Exception:
        pass
    query = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
    while not stop.is_set():
        try:
            out = subprocess.run(['nvidia-smi', f'--id={gpu_index}', f'--query-gpu={query}',
                                  '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
            parts = [p.strip() for p in out.strip().split(',')]
            if len(parts) >= 7:
                samples.append(parts[:7] + [time.perf_counter()])
        except Exception:
            pass
        stop.wait(0.2)


def _clocks_summary(samples, window=None):
    """ Median SM clock and the throttle reasons seen, over the samples inside `window` = (t0, t1) (the timed region;
    all samples when none fell inside it). """
    if window is not None:
        inside = [s for s in samples if window[0] <= s[-1] <= window[1]]
        if inside:
            samples = inside
    if not samples:
        return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unsampled'])
    sm = [float(s[0]) for s in samples if s[0].replace('.', '', 1).isdigit()]
    smax = [float(s[1]) for s in samples if s[1].replace('.', '', 1).isdigit()]
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    reasons = [n for i, n in enumerate(names) if any(str(s[3 + i]).lower().startswith('active') for s in samples)]
    return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(smax) if smax else None,
                reasons=reasons, samples=len(samples))


def _oracle_step(cfg, src_np, ref_np, src_tf, ref_tf, bands):
    """ The reference algorithm (oracle port) for `bands` bands; returns seconds. """
    from oracle import kernel_model_np as kmnp
    t0 = time.perf_counter()
    for b in bands:
        kmnp.fuse_band_blocks(src_np[b], src_tf, cfg['src_nodata'], ref_np[b], ref_tf, NAN, cfg['model'],
                              cfg['kernel_shape'], cfg['proc_crs'], False, cfg['r2_inpaint_thresh'])
    return time.perf_counter() - t0


def run_reference(args, cfg, rank):
    """ --impl reference: the reference's CPU path (oracle port) on host cores; rank 0 only. """
    if rank != 0:
        return
    import cv2
    import torch
    from homonim_b200.synthetic import make_pair
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    torch.set_num_threads(cores)
    # bounded sample: a crop for the source-resolution workload, so that K steps end within minutes.  All bands of the
    # sample run concurrently on a thread pool, as the reference does with its (band, block) jobs (fuse.py:396-408);
    # inside a band cv2 and the OpenMP resamplers use the remaining parallelism
    from concurrent.futures import ThreadPoolExecutor
    hp, wp = (cfg['hp'], cfg['wp']) if cfg['proc_crs'] == 'ref' else (min(cfg['hp'], 2048), min(cfg['wp'], 2048))
    n_bands = cfg['bands']
    src_ra, ref_ra = make_pair(hp, wp, cfg['ratio'], bands=n_bands, dtype=cfg['dtype'], mu=cfg['mu'], seed=2,
                               device='cpu', src_nodata=cfg['src_nodata'])
    src_np, ref_np = src_ra.to_host().array, ref_ra.to_host().array
    src_tf, ref_tf = tuple(src_ra.transform), tuple(ref_ra.transform)
    npix = src_np.size
    sample = (f'all {n_bands} bands concurrently (thread pool), {src_np.shape[1]}x{src_np.shape[2]} source pixels per '
              f'band and step' + ('' if cfg['proc_crs'] == 'ref' else ' (crop)'))
    workers = max(1, min(n_bands, cores))

    def one_step():
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=workers) as pool:
            list(pool.map(lambda b: _oracle_step(cfg, src_np, ref_np, src_tf, ref_tf, [b]), range(n_bands)))
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    times = [one_step() for _ in range(args.steps)]
    total = sum(times)
    value = npix * args.steps / total / 1e6
    line = {
        'impl': 'reference', 'metric': 'fit+apply Mpix/s', 'value': round(value, 3), 'unit': 'Mpix/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': round(1e3 * total / args.steps, 3), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['desc'], 'sample': sample},
        'cpu_baseline': {'value': round(value, 3), 'unit': 'Mpix/s', 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': round(value, 3), 'unit': 'Mpix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def _gpu_uuid(torch, index):
    """ UUID of CUDA device `index` (NVML enumerates physical devices; CUDA_VISIBLE_DEVICES may renumber them). """
    try:
        return str(torch.cuda.get_device_properties(index).uuid)
    except Exception:
        return None


def _peak():
    peaks_path = REPO / 'MEASURED_PEAKS.json'
    if peaks_path.exists():
        return float(json.loads(peaks_path.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def _ncu_traffic(workload, entry_point):
    """
    DRAM bytes (read + write) per launch of `entry_point` from the committed `ncu --set full` capture of this workload
    (profiles/ncu_traffic.json: {workload: {entry point: bytes}}); None when no capture is on record.
    """
    path = REPO / 'profiles' / 'ncu_traffic.json'
    if not path.exists():
        return None
    try:
        return json.loads(path.read_text()).get(workload, {}).get(entry_point)
    except Exception:
        return None


def run_sharded(args, cfg, rank, world, local_rank):
    """
    One raster split into row bands over the ranks (configuration C5a): strong scaling.  Every rank
    generates its own rows (seeded per rank; the reference rows are all-gathered into the replicated proc-grid
    reference), then per band: down-sample own rows -> all-gather the proc-grid plane -> fit (redundantly, < 1 % of the
    work) -> up-sample + apply own rows.
    """
    import torch
    import torch.distributed as dist
    from homonim_b200 import Affine, Model, RasterArray, RefSpaceModel, _native
    from homonim_b200.dist import RowBands, all_gather_rows, fuse_refspace_sharded
    from homonim_b200.kernel_model import KernelTimer
    from homonim_b200.synthetic import make_pair

    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29511')
    dist.init_process_group('nccl', device_id=device, rank=rank, world_size=world)
    lib = _native.lib()
    bands = RowBands.split(cfg['hp'], world)
    a, b = bands.band(rank)
    ratio, n_bands = cfg['ratio'], cfg['bands']
    # this rank's rows of the source, one band at a time (the float32 intermediates of a 60k-wide band are large)
    # (generated in chunks of 500 proc rows: a whole 60 000 x 60 000 float32 plane exceeds torch's 2^31-element limit
    # for bilinear interpolation, and its float32 intermediates would not fit beside the raster itself)
    src_planes, ref_rows = [], []
    chunk = 500
    tdtype = getattr(torch, cfg['dtype'])
    for band in range(n_bands):
        plane = torch.empty(((b - a) * ratio, cfg['wp'] * ratio), dtype=tdtype, device=device)
        rrows = torch.empty((b - a, cfg['wp']), dtype=torch.float32, device=device)
        for c0 in range(0, b - a, chunk):
            c1 = min(c0 + chunk, b - a)
            s_ra, r_ra = make_pair(c1 - c0, cfg['wp'], ratio, bands=1, dtype=cfg['dtype'],
                                   mu=cfg['mu'] * (1 + 0.15 * band), seed=50 + 1009 * rank + 31 * band + c0,
                                   device=device, src_nodata=cfg['src_nodata'], ref_pad=0)
            plane[c0 * ratio:c1 * ratio] = s_ra.array[0]
            rrows[c0:c1] = r_ra.array[0]
            crs, src_tf0 = s_ra.crs, s_ra.transform
            del s_ra, r_ra
        src_planes.append(plane)
        ref_rows.append(rrows)
        torch.cuda.empty_cache()
    # global grids: rank's source rows start at proc row a
    res = src_tf0.a
    x0, y0 = src_tf0.c, src_tf0.f
    ref_global_tf = Affine(res * ratio, 0.0, x0, 0.0, -res * ratio, y0)
    src_local_tf = Affine(res, 0.0, x0, 0.0, -res, y0 - a * ratio * res)
    ref_planes = [all_gather_rows(r, bands) for r in ref_rows]
    del ref_rows
    model = RefSpaceModel(Model(cfg['model']), cfg['kernel_shape'], r2_inpaint_thresh=cfg['r2_inpaint_thresh'])
    out = torch.empty_like(src_planes[0], dtype=torch.float32)          # one band of corrected rows, re-used
    npix_total = cfg['hp'] * cfg['wp'] * ratio * ratio * n_bands        # source band-pixels of the WHOLE raster

    def step():
        for band in range(n_bands):
            src_local = RasterArray(src_planes[band], crs, src_local_tf, nodata=cfg['src_nodata'])
            ref_ra = RasterArray(ref_planes[band], crs, ref_global_tf, nodata=NAN)
            fuse_refspace_sharded(model, src_local, ref_ra, bands, out=out)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    stop, samples = threading.Event(), []
    sampler = threading.Thread(target=_clock_sampler, args=(stop, samples, local_rank, _gpu_uuid(torch, local_rank)),
                               daemon=True)
    if rank == 0:
        sampler.start()
    lib.hb_reset_launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_region0 = time.perf_counter()
    start.record()
    for _ in range(args.steps):
        step()
    end.record()
    barrier()
    launches = lib.hb_launch_count()
    elapsed_ms = start.elapsed_time(end)
    with KernelTimer() as timer:
        step()
        kernel_ms = timer.results()
    t_region1 = time.perf_counter()
    stop.set()
    if rank == 0:
        sampler.join(timeout=2)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = npix_total * args.steps / (elapsed_ms * 1e-3) / 1e6
    peak_gbs, peak_src = _peak()
    b_in = src_planes[0].element_size()
    local_px = src_planes[0].numel()
    alg = {'hb_upsample_apply': local_px * (b_in + 4), 'hb_downsample_average': local_px * b_in}
    per_kernel = {k: {'launches': len(v), 'ms_avg': round(sum(v) / len(v), 4)} for k, v in kernel_ms.items()}
    for k, d in per_kernel.items():
        if k in alg:
            d['gbs'] = round(alg[k] / (d['ms_avg'] * 1e-3) / 1e9, 1)
    dominant = max((k for k in kernel_ms if k in alg), key=lambda k: sum(kernel_ms[k]))
    achieved = alg[dominant] / (sum(kernel_ms[dominant]) / len(kernel_ms[dominant]) * 1e-3) / 1e9
    if rank == 0:
        line = {
            'metric': 'fit+apply Mpix/s', 'value': round(value, 1), 'unit': 'Mpix/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(elapsed_ms / args.steps, 4),
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': cfg['desc'], 'kernel_model': cfg['model'], 'kernel_shape': list(cfg['kernel_shape']),
                       'proc_crs': cfg['proc_crs'], 'bands': n_bands, 'src_dtype': cfg['dtype'],
                       'pixels_per_step': int(npix_total),
                       'sharding': f'row bands of {cfg["hp"]} proc rows over {world} rank(s); per band one all-gather '
                                   f'of the {cfg["hp"]}x{cfg["wp"]} float32 proc-grid plane',
                       'l2': 'inputs larger than L2 (no flush needed)'},
            'clocks': _clocks_summary(samples, (t_region0, t_region1)), 'e2e': None, 'gpu_launches': int(launches),
            'roofline': {'bound': 'hbm', 'kernel': dominant, 'achieved': round(achieved, 1), 'peak': peak_gbs,
                         'peak_source': peak_src, 'unit': 'GB/s', 'frac': round(achieved / peak_gbs, 4),
                         'traffic': None, 'algorithmic_bytes_per_launch': int(alg[dominant]), 'kernels': per_kernel},
            'cpu_baseline': None,
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        # all host threads (torchrun exports OMP_NUM_THREADS=1; the OpenMP runtime reads it when the oracle's C
        # restatement is first loaded, so override it before anything is imported)
        os.environ['OMP_NUM_THREADS'] = str(os.cpu_count() or 1)
        run_reference(args, cfg, rank)
        return
    if world == 1:
        os.environ.setdefault('OMP_NUM_THREADS', str(os.cpu_count() or 1))
    if cfg.get('sharded'):
        run_sharded(args, cfg, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist
    from homonim_b200 import Model, ProcCrs, RasterArray, RasterFuse, _native
    from homonim_b200.kernel_model import KernelTimer
    from homonim_b200.synthetic import make_pair

    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    lib = _native.lib()

    # ---- inputs: every rank corrects its own source image (seeded per rank), resident in HBM ------------------------
    src_ra, ref_ra = make_pair(cfg['hp'], cfg['wp'], cfg['ratio'], bands=cfg['bands'], dtype=cfg['dtype'],
                               mu=cfg['mu'], seed=2 + rank, device=device, src_nodata=cfg['src_nodata'])
    npix = src_ra.array.numel()                      # source band-pixels per step
    model_config = dict(r2_inpaint_thresh=cfg['r2_inpaint_thresh'])
    fuse = RasterFuse(src_ra, ref_ra, proc_crs=ProcCrs(cfg['proc_crs']))
    fuse.open()

    def step(f=fuse, serial=False):
        # serial=True: one CUDA stream (bands back to back) -- used only for the per-kernel roofline timing pass
        return f.process(model=Model(cfg['model']), kernel_shape=cfg['kernel_shape'], model_config=model_config,
                         block_config=dict(threads=1) if serial else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: device-resident inputs -----------------------------------------------------------------------
    # clocks are sampled on rank 0 only (nvidia-smi is not free: 8 ranks polling it would compete with the timed loop)
    stop, samples = threading.Event(), []
    sampler = threading.Thread(target=_clock_sampler, args=(stop, samples, local_rank, _gpu_uuid(torch, local_rank)),
                               daemon=True)
    if rank == 0:
        sampler.start()
    lib.hb_reset_launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_region0 = time.perf_counter()
    start.record()
    for _ in range(args.steps):
        step()
    end.record()
    barrier()
    launches = lib.hb_launch_count()
    elapsed_ms = start.elapsed_time(end)
    # per-launch kernel durations for the roofline: the same K steps again with the bands on ONE stream (concurrent
    # bands would overlap kernels and inflate each other's event-to-event durations)
    with KernelTimer() as timer:
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            step(serial=True)
        s1.record()
        kernel_ms = timer.results()
    serial_ms = s0.elapsed_time(s1)
    t_region1 = time.perf_counter()          # (clock window: the timed steps and the same steps serialised for the roofline)
    stop.set()
    if rank == 0:
        sampler.join(timeout=2)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = npix * world * args.steps / (elapsed_ms * 1e-3) / 1e6

    # ---- end to end: the same public call with HOST rasters (pinned), H2D + D2H inside the timed region ---------------
    e2e = None
    if not args.no_e2e:
        src_host = torch.empty(src_ra.array.shape, dtype=src_ra.array.dtype, pin_memory=True)
        src_host.copy_(src_ra.array)
        ref_host = torch.empty(ref_ra.array.shape, dtype=ref_ra.array.dtype, pin_memory=True)
        ref_host.copy_(ref_ra.array)
        out_host = torch.empty(src_ra.array.shape, dtype=torch.float32, pin_memory=True)
        h2d = src_host.numel() * src_host.element_size() + ref_host.numel() * ref_host.element_size()
        d2h = out_host.numel() * out_host.element_size()

        def e2e_step():
            # the public call with HOST rasters (pinned CPU tensors): per band, host -> device copy of the source,
            # fit + apply, device -> host copy of the corrected band into `corr_out`; process() returns when the
            # corrected image is in host memory
            s = RasterArray(src_host, src_ra.crs, src_ra.transform, nodata=src_ra.nodata)
            r = RasterArray(ref_host, ref_ra.crs, ref_ra.transform, nodata=ref_ra.nodata)
            with RasterFuse(s, r, proc_crs=ProcCrs(cfg['proc_crs'])) as f:
                f.process(model=Model(cfg['model']), kernel_shape=cfg['kernel_shape'], model_config=model_config,
                          corr_out=out_host)

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e_steps = max(2, min(args.steps, 5))
        for _ in range(e_steps):
            e2e_step()
        barrier()
        e_ms = (time.perf_counter() - t0) * 1e3
        te = torch.tensor([e_ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {'value': round(npix * world * e_steps / (float(te.item()) * 1e-3) / 1e6, 2), 'unit': 'Mpix/s',
               'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h), 'steps': e_steps}
        del src_host, ref_host, out_host

    # ---- roofline of the dominant kernel (per-launch CUDA-event durations from the timed region) ---------------------
    peak_gbs, peak_src = _peak()
    b_in = src_ra.array.element_size()
    band_px = npix // cfg['bands']
    proc_px = (cfg['hp'] * cfg['wp']) if cfg['proc_crs'] == 'ref' else band_px
    want_r2 = cfg['model'] == 'gain-offset' and cfg['r2_inpaint_thresh'] is not None
    alg_bytes = {                                    # algorithmic bytes per launch
        'hb_upsample_apply': band_px * (b_in + 4) + 8 * proc_px,
        'hb_downsample_average': band_px * b_in + 4 * proc_px,
        'hb_fit_same_grid': proc_px * (8 + 4 * (3 if want_r2 else 2)),
        'hb_apply_same_grid': proc_px * 16,
        'hb_resample_up': proc_px * 4 + (cfg['hp'] * cfg['wp']) * 4,
    }
    per_kernel = {k: {'launches': len(v), 'ms_total': round(sum(v), 3), 'ms_avg': round(sum(v) / len(v), 4)}
                  for k, v in kernel_ms.items()}
    dominant = max((k for k in kernel_ms if k in alg_bytes), key=lambda k: sum(kernel_ms[k]))
    avg_s = sum(kernel_ms[dominant]) / len(kernel_ms[dominant]) * 1e-3
    achieved = alg_bytes[dominant] / avg_s / 1e9
    for k, d in per_kernel.items():
        if k in alg_bytes:
            d['gbs'] = round(alg_bytes[k] / (d['ms_avg'] * 1e-3) / 1e9, 1)
    step_bytes_per_px = (2 * b_in + 4) if cfg['proc_crs'] == 'ref' else 12      # DESIGN.md section 4 / SURVEY.md 8d
    roofline = {'bound': 'hbm', 'kernel': dominant, 'achieved': round(achieved, 1), 'peak': peak_gbs,
                'peak_source': peak_src, 'unit': 'GB/s', 'frac': round(achieved / peak_gbs, 4),
                'frac_of_nominal_8000_gbs': round(achieved / 8000.0, 4),
                'traffic': _ncu_traffic(args.workload, dominant),
                'algorithmic_bytes_per_launch': int(alg_bytes[dominant]),
                'share_of_step': round(sum(kernel_ms[dominant]) / serial_ms, 3),
                'whole_step': {'algorithmic_bytes_per_pixel': step_bytes_per_px,
                               'achieved_gbs': round(value * 1e6 * step_bytes_per_px / 1e9 / world, 1),
                               'frac': round(value * 1e6 * step_bytes_per_px / 1e9 / world / peak_gbs, 4),
                               'note': 'per GPU: value x algorithmic bytes per source band-pixel of the whole fit + '
                                       'apply step, against the same peak'},
                'timing': 'CUDA events around every launch, K steps with the bands serialised on one stream '
                          f'({round(serial_ms / args.steps, 4)} ms/step); `value` runs the bands on concurrent streams',
                'kernels': per_kernel}

    # ---- CPU baseline: the oracle port on a bounded sample (rank 0, N = 1 only) ---------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import cv2
        cores = os.cpu_count() or 1
        cv2.setNumThreads(cores)
        if cfg['proc_crs'] == 'ref':
            src_np, ref_np = src_ra.array[:1].cpu().numpy(), ref_ra.array[:1].cpu().numpy()
            src_tf, ref_tf = tuple(src_ra.transform), tuple(ref_ra.transform)
            sample = f'band 1 of {cfg["bands"]}, full {src_np.shape[1]}x{src_np.shape[2]} source band'
        else:
            c = 4096
            src_np = src_ra.array[:1, :c, :c].cpu().numpy()
            ref_np = ref_ra.array[:1, :c // cfg['ratio'] + 2, :c // cfg['ratio'] + 2].cpu().numpy()
            src_tf, ref_tf = tuple(src_ra.transform), tuple(ref_ra.transform)
            sample = f'band 1 of {cfg["bands"]}, {c}x{c} crop of the source band'
        _oracle_step(cfg, src_np, ref_np, src_tf, ref_tf, [0])      # warm-up (page-in, cv2 thread pool)
        secs = min(_oracle_step(cfg, src_np, ref_np, src_tf, ref_tf, [0]) for _ in range(2))
        cpu_baseline = {'value': round(src_np[0].size / secs / 1e6, 2), 'unit': 'Mpix/s', 'cores': cores,
                        'kind': 'port', 'sample': sample}

    if rank == 0:
        line = {
            'metric': 'fit+apply Mpix/s', 'value': round(value, 1), 'unit': 'Mpix/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(elapsed_ms / args.steps, 4),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': cfg['desc'], 'kernel_model': cfg['model'], 'kernel_shape': list(cfg['kernel_shape']),
                       'proc_crs': cfg['proc_crs'], 'bands': cfg['bands'], 'src_dtype': cfg['dtype'],
                       'pixels_per_step_per_gpu': int(npix), 'sharding': 'one source image per GPU, no collective',
                       'l2': 'inputs larger than L2 (no flush needed)'},
            'clocks': _clocks_summary(samples, (t_region0, t_region1)), 'e2e': e2e, 'gpu_launches': int(launches),
            'roofline': roofline,
            'cpu_baseline': cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""
bench.py -- fit + apply throughput of the kernel-model hot path (BASELINE.json metric) on N B200 GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload c2|c2-gain|c3|c4|c5a|c5b|tiny] [--no-row-band] [--no-e2e] [--no-cpu-baseline]

One "step" is one pass of the hot path over one synthetic source image: RasterFuse.process() = per band
RefSpaceModel/SrcSpaceModel .fit() + .apply().  Default workload: BASELINE.json configs[1] (C2) -- a 4-band uint16
10 000 x 10 000 aerial image against a 10 m reference (20x coarser), gain-offset 15x15 with r2_inpaint_thresh = 0.25,
proc_crs = ref.  With N > 1 every rank corrects its own source image (the batch-mosaic regime, no data-path
collective): weak scaling, `value` = pixels of all ranks / max-over-ranks time.

The same invocation ALSO measures the one-raster regime (BASELINE.json configs[4]) and carries it in
`row_band`: ONE 60 000 x 60 000 4-band float32 raster split into row bands over the N ranks (strong scaling; block
statistics merged over the ranks, P2P halos of the down-sampled rows, homonim_b200/dist.py), with -- for N > 1 -- the
single-GPU time of the same raster measured on rank 0 in the same process (`row_band.n1`, `efficiency_vs_n1`) and a
sharded-vs-unsharded comparison of row stripes straddling shard boundaries (`row_band.parity`).  `--workload c5a|c5b`
makes that regime the headline `value` instead.

The JSON line carries: `value` (device-resident inputs, CUDA-event timed), `e2e` (the same call with HOST buffers:
pinned host -> device copies of the inputs and device -> host copy of the corrected image inside the timed region,
plus the bare copy bandwidths), `roofline` for the dominant kernel (per-launch CUDA-event durations collected live),
`parity` (N = 1: the GPU result of the sampled band against the reference's own result on the same inputs: masks,
max relative errors, SURVEY.md 8d metric), `cpu_baseline` (the reference on a bounded sample, host cores), and
`clocks` sampled through NVML during the timed region.

`--impl reference` times the reference's CPU implementation of the same path on the host cores: the UNMODIFIED
reference package (installed under baseline/_ref by __graft_entry__.build(); its numpy + cv2 kernel-model code, with
its two GDAL calls served by oracle/gdal_restate -- rasterio/GDAL are not installable here) in the reference's own
default mode (blocks of max_block_mem = 100 MB on a thread pool of all cores, fuse.py:396-408) and in one-block mode;
the oracle port stands in when baseline/_ref is absent (`kind: "port"`).
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
    # C1 (BASELINE.json configs[0]): the reference's own test images (real data, mis-aligned grids), committed as a fixture
    'c1': dict(fixture='docs_cli_ngi1', hp=711, wp=403, ratio=2, bands=3, dtype='uint8', mu=100.0, src_nodata=0.0,
               model='gain-blk-offset', kernel_shape=(5, 5), r2_inpaint_thresh=0.25, proc_crs='ref',
               src_shape=(1421, 805),
               desc='C1: the reference\'s test images ngi_rgb_byte_1.tif (3-band uint8 1421x805, 5 m) + sentinel2_b432_byte.tif '
                    '(10 m, grids mis-aligned by a fraction of a pixel), gain-blk-offset 5x5, proc_crs=ref'),
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
                     'gain-blk-offset 15x15, proc_crs=ref'),
    'c5b': dict(hp=60000, wp=60000, ratio=1, bands=4, dtype='float32', mu=0.3, src_nodata=NAN,
                model='gain-blk-offset', kernel_shape=(15, 15), r2_inpaint_thresh=0.25, proc_crs='src', sharded=True,
                desc='C5b: ONE synthetic 4-band float32 60000x60000 raster with the reference on the same grid, sharded '
                     'as row bands over the GPUs, gain-blk-offset 15x15 (same-grid fit + apply)'),
    'tiny-band': dict(hp=96, wp=80, ratio=8, bands=2, dtype='float32', mu=0.3, src_nodata=NAN,
                      model='gain-blk-offset', kernel_shape=(5, 5), r2_inpaint_thresh=0.25, proc_crs='ref',
                      sharded=True, desc='tiny-band: 2-band float32 768x640 raster as row bands (script test only)'),
}
ROW_BAND_CHUNK = 125          # proc rows per generation chunk of the sharded rasters (independent of the rank count)


def _data(cfg):
    return ('real: the reference\'s own test images (tests/golden fixture)' if cfg.get('fixture') else 'synthetic')


def _small(cfg):
    """ Workloads whose inputs fit the 126 MB L2: timed step by step with an L2 flush in between. """
    hs, ws = cfg.get('src_shape') or (cfg['hp'] * cfg['ratio'], cfg['wp'] * cfg['ratio'])
    return hs * ws * cfg['bands'] * 4 < (256 << 20)


def _config(cfg):
    """ The `config` object of the JSON line -- identical for both arms (`--impl b200` / `--impl reference`). """
    hs, ws = cfg.get('src_shape') or (cfg['hp'] * cfg['ratio'], cfg['wp'] * cfg['ratio'])
    return {'workload': cfg['desc'], 'kernel_model': cfg['model'], 'kernel_shape': list(cfg['kernel_shape']),
            'r2_inpaint_thresh': cfg['r2_inpaint_thresh'], 'proc_crs': cfg['proc_crs'], 'bands': cfg['bands'],
            'src_dtype': cfg['dtype'], 'src_shape': [hs, ws], 'pixels_per_step_per_image': int(hs * ws * cfg['bands']),
            'l2': ('inputs fit in L2: a 512 MB buffer is written between timed steps (each step timed on its own)'
                   if _small(cfg) else 'inputs larger than L2 (no flush needed)')}


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
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
    (profiles/ncu_traffic.json: {workload: {entry point: bytes}}) -- a STATIC figure read from the repository, not
    measured in this run; None when no capture is on record.
    """
    path = REPO / 'profiles' / 'ncu_traffic.json'
    if not path.exists():
        return None
    try:
        return json.loads(path.read_text()).get(workload, {}).get(entry_point)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------------------
# the reference on the host cores (oracle / test infrastructure: only the checker and the reported CPU baseline)
# ---------------------------------------------------------------------------------------------------------------------
class CpuReference:
    """ The reference's CPU implementation of one band's fit + apply: the unmodified package when baseline/_ref (or
    /root/reference) is present, else the oracle port. """

    def __init__(self):
        from oracle import ref_runner
        self.runner = ref_runner
        self.ns = ref_runner.load()
        self.kind = 'reference' if self.ns is not None else 'port'
        self.where = self.ns.root if self.ns is not None else 'oracle/kernel_model_np.py'

    def one_block(self, cfg, src, src_tf, ref, ref_tf, find_r2=False):
        """ (params, corr) of one band as ONE block. """
        if self.ns is not None:
            return self.runner.fuse_band(self.ns, src, src_tf, cfg['src_nodata'], ref, ref_tf, cfg['model'],
                                         cfg['kernel_shape'], cfg['proc_crs'], cfg['r2_inpaint_thresh'], find_r2=find_r2)
        from oracle import kernel_model_np as kmnp
        params, _, corr = kmnp.fuse_band_blocks(src, src_tf, cfg['src_nodata'], ref, ref_tf, NAN, cfg['model'],
                                                cfg['kernel_shape'], cfg['proc_crs'], find_r2, cfg['r2_inpaint_thresh'])
        return params, corr

    def blocked(self, cfg, src_bands, src_tf, ref_bands, ref_tf, threads, max_block_mem=100.0):
        """ All bands through the reference's block grid on ONE thread pool; returns (seconds, blocks). """
        from concurrent.futures import ThreadPoolExecutor
        t0 = time.perf_counter()
        n_blocks = 0
        with ThreadPoolExecutor(max_workers=threads) as pool:
            futures = []
            for b in range(len(src_bands)):
                _, futs = self.runner.fuse_band_blocked(self.ns, src_bands[b], src_tf, cfg['src_nodata'], ref_bands[b],
                                                        ref_tf, cfg['model'], cfg['kernel_shape'], cfg['proc_crs'],
                                                        cfg['r2_inpaint_thresh'], max_block_mem=max_block_mem,
                                                        executor=pool)
                futures += futs
            for f in futures:
                f.result()
            n_blocks = len(futures)
        return time.perf_counter() - t0, n_blocks


def _fixture_pair(cfg, device):
    """ The C1 pair from tests/golden (the reference's own test images, read once from its GeoTIFFs by
    oracle/make_golden_docs.py): uint8 source with nodata 0, reference widened to float32 (it has no nodata). """
    import torch
    from homonim_b200 import CRS, Affine, RasterArray
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tests', 'golden')
    meta = json.load(open(os.path.join(root, cfg['fixture'] + '.json')))
    with np.load(os.path.join(root, cfg['fixture'] + '.npz')) as data:
        src, ref = data['src'], data['s2'].astype('float32')
    crs = CRS.from_epsg(32735)
    src_ra = RasterArray(torch.from_numpy(src).to(device), crs, Affine(*meta['src_transform'][:6]), nodata=cfg['src_nodata'])
    ref_ra = RasterArray(torch.from_numpy(ref).to(device), crs, Affine(*meta['s2_transform'][:6]), nodata=float('nan'))
    return src_ra, ref_ra


def _cpu_sample(cfg, src_ra, ref_ra, n_bands):
    """ Host arrays of the bounded CPU sample: whole bands for the proc_crs = ref workloads, a crop for the workloads
    that fit at source resolution (the full 20 000^2 float32 planes would take minutes per band on the host). """
    if cfg['proc_crs'] == 'ref':
        src_np, ref_np = src_ra.array[:n_bands].cpu().numpy(), ref_ra.array[:n_bands].cpu().numpy()
        note = f'{n_bands} of {cfg["bands"]} band(s), full {src_np.shape[1]}x{src_np.shape[2]} source band'
    else:
        c = min(4096, src_ra.array.shape[-1])
        rc = c // cfg['ratio'] + 2
        src_np = src_ra.array[:n_bands, :c, :c].cpu().numpy()
        ref_np = ref_ra.array[:n_bands, :rc, :rc].cpu().numpy()
        note = f'{n_bands} of {cfg["bands"]} band(s), {c}x{c} crop of the source band'
    return src_np, ref_np, tuple(src_ra.transform), tuple(ref_ra.transform), note


def run_reference(args, cfg, rank):
    """ --impl reference: the reference's CPU path on host cores; rank 0 only. """
    if rank != 0:
        return
    import cv2
    import torch
    from homonim_b200.synthetic import make_pair
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    torch.set_num_threads(cores)
    cpu = CpuReference()
    if cfg.get('sharded'):
        # one row stripe of the big raster per step (the CPU arm can neither hold nor finish 57.6 GB per step)
        hp, wp = (min(cfg['hp'], ROW_BAND_CHUNK), cfg['wp']) if cfg['ratio'] > 1 else (min(cfg['hp'], 1024), cfg['wp'])
    else:
        hp, wp = (cfg['hp'], cfg['wp']) if cfg['proc_crs'] == 'ref' else (min(cfg['hp'], 2048), min(cfg['wp'], 2048))
    if cfg.get('fixture'):
        src_ra, ref_ra = _fixture_pair(cfg, 'cpu')
    else:
        src_ra, ref_ra = make_pair(hp, wp, cfg['ratio'], bands=cfg['bands'], dtype=cfg['dtype'], mu=cfg['mu'], seed=2,
                                   device='cpu', src_nodata=cfg['src_nodata'])
    n_bands = cfg['bands']
    src_np, ref_np = src_ra.to_host().array, ref_ra.to_host().array
    src_tf, ref_tf = tuple(src_ra.transform), tuple(ref_ra.transform)
    npix = src_np.size
    crop = '' if (cfg['proc_crs'] == 'ref' and not cfg.get('sharded')) else ' (a crop / stripe of the workload)'
    sample = f'all {n_bands} bands, {src_np.shape[1]}x{src_np.shape[2]} source pixels per band and step{crop}'
    from concurrent.futures import ThreadPoolExecutor

    def one_block_step():
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=max(1, min(n_bands, cores))) as pool:
            list(pool.map(lambda b: cpu.one_block(cfg, src_np[b], src_tf, ref_np[b], ref_tf), range(n_bands)))
        return time.perf_counter() - t0

    def blocked_step():
        return cpu.blocked(cfg, src_np, src_tf, ref_np, ref_tf, threads=cores)[0]

    modes = {'one_block': one_block_step}
    if cpu.kind == 'reference':
        modes['blocked'] = blocked_step
    # bounded sample: the whole `--steps K --warmup W` run (both modes) is to end within a few minutes, so after one
    # calibration step on all bands the step is cut down to as many bands as fit a ~150 s budget (throughput per band is
    # what is reported; the bands are independent in the reference too)
    t_cal = one_block_step()
    budget_s = float(os.environ.get('HB_REFERENCE_BUDGET_S', '150'))
    projected = t_cal * len(modes) * (args.steps + args.warmup)
    if projected > budget_s and n_bands > 1:
        n_bands = max(1, min(n_bands, int(n_bands * budget_s / projected)))
        src_np, ref_np = src_np[:n_bands], ref_np[:n_bands]
        npix = src_np.size
        sample = (f'{n_bands} of {cfg["bands"]} bands (bounded sample: a calibration step on all bands took {t_cal:.1f} s), '
                  f'{src_np.shape[1]}x{src_np.shape[2]} source pixels per band and step{crop}')
    results = {}
    for name, fn in modes.items():
        for _ in range(args.warmup):
            fn()
        times = [fn() for _ in range(args.steps)]
        results[name] = {'value': round(npix * args.steps / sum(times) / 1e6, 3), 'unit': 'Mpix/s',
                         'ms_per_step': round(1e3 * sum(times) / args.steps, 3)}
    # headline of this arm: the reference's own default mode (threads = all cores, max_block_mem = 100 MB) when the
    # unmodified reference is present; the one-block run otherwise
    head = 'blocked' if 'blocked' in results else 'one_block'
    value, ms = results[head]['value'], results[head]['ms_per_step']
    results['one_block']['mode'] = 'one block per band, bands concurrently on a thread pool (what the GPU path computes)'
    if 'blocked' in results:
        _, n_blocks = cpu.blocked(cfg, src_np[:1], src_tf, ref_np[:1], ref_tf, threads=cores)
        results['blocked']['mode'] = (f'the reference\'s default: max_block_mem=100 MB -> {n_blocks} block(s) per band, '
                                      f'overlap ceil(k/2), ThreadPoolExecutor(max_workers={cores})')
    line = {
        'impl': 'reference', 'metric': 'fit+apply Mpix/s', 'value': value, 'unit': 'Mpix/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'strong' if cfg.get('sharded') else 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': _data(cfg), 'config': _config(cfg),
        'cpu_baseline': {'value': value, 'unit': 'Mpix/s', 'cores': cores, 'kind': cpu.kind, 'sample': sample,
                         'headline_mode': head, 'modes': results, 'implementation': cpu.where,
                         'gdal': 'rasterio.warp.reproject / rasterio.fill.fillnodata served by oracle/gdal_restate '
                                 '(OpenMP C restatement; rasterio / GDAL are not installable in this image)'},
        'e2e': {'value': value, 'unit': 'Mpix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# parity of the GPU result against the reference's result (SURVEY.md 8d metric)
# ---------------------------------------------------------------------------------------------------------------------
def _rel_err(actual, expected, floor, exclude):
    ok = np.isfinite(expected) & ~exclude
    if not ok.any():
        return 0.0
    a, e = actual[ok].astype('float64'), expected[ok].astype('float64')
    fl = np.broadcast_to(np.asarray(floor, dtype='float64'), expected.shape)[ok]
    return float(np.max(np.abs(a - e) / np.maximum(np.abs(e), fl)))


def parity_metrics(got_params, got_corr, exp_params, exp_corr, src_mean, param_to_corr=None, offset_at_corr=None,
                   kernel_px=None):
    """
    SURVEY.md 8(d): masks `isnan(out) == isnan(reference)` exactly; gain / corrected-pixel error relative to
    max(|x_ref|, 1e-3 x band mean); offset error relative to max(|offset|, |gain| x mean(src)); R2 absolute.  Pixels
    where the reference's own solve is ill-conditioned (|gain| > 50 x the band median: a denominator crossing zero
    turns rounding noise into the value) are counted in `excluded_px` and left out of the maxima --
    never out of the mask comparison.  ``param_to_corr``: maps a boolean parameter-grid mask to the corrected-image grid
    (the pixels whose cubic-spline taps touch those parameters), so that the same pixels are left out there.
    ``offset_at_corr``: the reference's offset parameter looked up at every corrected pixel; corrected pixels where the
    reference's own float32 expression gain * src + offset (kernel_model.py:461) cancels by more than 2^7 (|corr| < |offset| /
    128: the value is then the rounding residue of two ~100x larger float32 terms) are counted in `cancellation_px` and
    likewise left out of `max_rel_err_corr` -- their worst error is reported separately, relative to the same floor.
    ``kernel_px`` = 1 (a 1 x 1 kernel): R2 = 1 - ss_res / ss_tot with ss_tot = ref^2 - ref^2 / 1, i.e. BOTH terms are the
    rounding residue of the reference's own float32 expression (kernel_model.py:280-302) -- the band carries no
    information; it is reported (`r2_abs`, `r2_bit_identical_frac`) and flagged `r2_degenerate`, not held to 1e-4.
    """
    n = min(got_params.shape[0], exp_params.shape[0])
    masks = bool(np.array_equal(np.isnan(got_corr), np.isnan(exp_corr)) and
                 all(np.array_equal(np.isnan(got_params[b]), np.isnan(exp_params[b])) for b in range(n)))
    infs_ok = all(np.array_equal(got_params[b][np.isinf(exp_params[b])], exp_params[b][np.isinf(exp_params[b])])
                  for b in range(n))
    gain_e = exp_params[0]
    fin = np.isfinite(gain_e)
    med = float(np.median(np.abs(gain_e[fin]))) if fin.any() else 1.0
    with np.errstate(invalid='ignore'):
        bad = np.abs(gain_e) > 50 * max(med, 1e-30)
    g_floor = 1e-3 * float(np.mean(np.abs(gain_e[fin & ~bad]))) if (fin & ~bad).any() else 1.0
    o_floor = np.maximum(np.abs(np.nan_to_num(gain_e, nan=0.0, posinf=0.0, neginf=0.0)) * abs(src_mean), 1e-30)
    out = {'masks_identical': masks, 'infinities_identical': bool(infs_ok),
           'max_rel_err_gain': _rel_err(got_params[0], gain_e, g_floor, bad),
           'max_rel_err_offset': _rel_err(got_params[1], exp_params[1], o_floor, bad)}
    # corrected pixels: the ill-conditioned parameter pixels are mapped to the source grid by the caller through
    # `excluded_corr` when grids differ; on the corrected image the same |value| criterion is applied directly
    fin_c = np.isfinite(exp_corr)
    med_c = float(np.median(np.abs(exp_corr[fin_c]))) if fin_c.any() else 1.0
    with np.errstate(invalid='ignore'):
        bad_c = np.abs(exp_corr) > 50 * max(med_c, 1e-30)
    if bad.any():
        bad_c |= param_to_corr(bad) if param_to_corr is not None else bad
    c_floor = 1e-3 * float(np.mean(np.abs(exp_corr[fin_c & ~bad_c]))) if (fin_c & ~bad_c).any() else 1.0
    cancel = np.zeros(exp_corr.shape, bool)
    if offset_at_corr is not None:
        with np.errstate(invalid='ignore'):
            cancel = fin_c & (np.abs(exp_corr) * 128.0 < np.abs(offset_at_corr))
        out['cancellation_px'] = int(cancel.sum())
        out['max_rel_err_corr_cancellation_px'] = float(f'{_rel_err(got_corr, exp_corr, c_floor, ~cancel | bad_c):.3e}')
    out['max_rel_err_corr'] = _rel_err(got_corr, exp_corr, c_floor, bad_c | cancel)
    if n > 2:
        fin2 = np.isfinite(exp_params[2]) & ~bad
        out['r2_abs'] = float(np.max(np.abs(got_params[2][fin2].astype('float64') - exp_params[2][fin2]))) \
            if fin2.any() else 0.0
        same_r2 = (got_params[2] == exp_params[2]) | (np.isnan(got_params[2]) & np.isnan(exp_params[2]))
        out['r2_bit_identical_frac'] = round(float(same_r2.mean()), 6)
        if kernel_px == 1:
            out['r2_degenerate'] = '1 x 1 kernel: ss_tot is identically 0 up to float32 rounding; R2 not held to 1e-4'
    out['excluded_px'] = int(bad.sum())
    out['excluded_corr_px'] = int(bad_c.sum())
    same = (got_corr == exp_corr) | (np.isnan(got_corr) & np.isnan(exp_corr))
    out['corr_bit_identical_frac'] = round(float(same.mean()), 6)
    out['tolerance'] = 1e-4
    out['within_tolerance'] = bool(masks and infs_ok and out['max_rel_err_gain'] <= 1e-4 and
                                   out['max_rel_err_offset'] <= 1e-4 and out['max_rel_err_corr'] <= 1e-4 and
                                   (out.get('r2_abs', 0.0) <= 1e-4 or 'r2_degenerate' in out))
    for k in ('max_rel_err_gain', 'max_rel_err_offset', 'max_rel_err_corr', 'r2_abs'):
        if k in out:
            out[k] = float(f'{out[k]:.3e}')
    return out


def measure_parity_and_cpu(args, cfg, src_ra, ref_ra):
    """ N = 1, rank 0: the reference on a bounded sample (timed: `cpu_baseline`) and the GPU result of the same sample
    against it (`parity`). """
    import cv2
    import torch
    from homonim_b200 import Model, ProcCrs, RasterArray, RasterFuse
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    cpu = CpuReference()
    src_np, ref_np, src_tf, ref_tf, note = _cpu_sample(cfg, src_ra, ref_ra, 1)
    cpu.one_block(cfg, src_np[0], src_tf, ref_np[0], ref_tf)          # warm-up (page-in, cv2 thread pool)
    secs = []
    for _ in range(2):
        t0 = time.perf_counter()
        cpu.one_block(cfg, src_np[0], src_tf, ref_np[0], ref_tf)
        secs.append(time.perf_counter() - t0)
    cpu_baseline = {'value': round(src_np[0].size / min(secs) / 1e6, 2), 'unit': 'Mpix/s', 'cores': cores,
                    'kind': cpu.kind, 'sample': note + ', one block (what the GPU path computes)',
                    'implementation': cpu.where}
    # parity: the reference's result WITH the R2 band, against RasterFuse.process(param_filename=...) on the same sample
    exp_params, exp_corr = cpu.one_block(cfg, src_np[0], src_tf, ref_np[0], ref_tf, find_r2=True)
    dev = src_ra.array.device
    s = RasterArray(torch.from_numpy(src_np).to(dev), src_ra.crs, src_ra.transform, nodata=src_ra.nodata)
    r = RasterArray(torch.from_numpy(ref_np).to(dev), ref_ra.crs, ref_ra.transform, nodata=ref_ra.nodata)
    with RasterFuse(s, r, proc_crs=ProcCrs(cfg['proc_crs'])) as f:
        corr_ra, param_ra = f.process(model=Model(cfg['model']), kernel_shape=cfg['kernel_shape'],
                                      model_config=dict(r2_inpaint_thresh=cfg['r2_inpaint_thresh']),
                                      param_filename='params')
    got_corr = corr_ra.array[0].cpu().numpy()
    got_params = param_ra.array.cpu().numpy()            # [n_params * 1 band, h, w]
    if got_params.shape[1:] != exp_params.shape[1:]:
        raise RuntimeError(f'parameter grids differ: {got_params.shape} vs {exp_params.shape}')
    valid = src_np[0][~np.isnan(src_np[0].astype('float32'))] if np.isnan(cfg['src_nodata']) else \
        src_np[0][src_np[0] != cfg['src_nodata']]
    param_to_corr, offset_at_corr = None, exp_params[1]
    if cfg['proc_crs'] == 'ref':
        st, pt = src_ra.transform, param_ra.transform
        hs, ws = got_corr.shape
        rows = np.floor(((st.f + (np.arange(hs) + 0.5) * st.e) - pt.f) / pt.e).astype(int).clip(0, exp_params.shape[1] - 1)
        cols = np.floor(((st.c + (np.arange(ws) + 0.5) * st.a) - pt.c) / pt.a).astype(int).clip(0, exp_params.shape[2] - 1)

        def param_to_corr(mask):
            # a parameter pixel reaches the source pixels under the 4 x 4 spline taps around it: dilate by 2, look up
            grown = cv2.dilate(mask.astype('uint8'), np.ones((5, 5), 'uint8')) > 0
            return grown[np.ix_(rows, cols)]
        offset_at_corr = exp_params[1][np.ix_(rows, cols)]
    parity = parity_metrics(got_params, got_corr, exp_params, exp_corr, float(valid.astype('float64').mean()),
                            param_to_corr, offset_at_corr, kernel_px=int(np.prod(cfg['kernel_shape'])))
    parity['sample'] = note
    parity['against'] = f'{cpu.kind} ({cpu.where})'
    if cfg['model'] == 'gain-offset' and cfg['r2_inpaint_thresh'] is not None:
        parity['note'] = ('in-painting parity unpinned: GDALFillNodata is compared with a restatement of its published '
                          'algorithm, not with GDAL (DESIGN.md section 2)')
    return cpu_baseline, parity


# ---------------------------------------------------------------------------------------------------------------------
# one raster as row bands (C5a / C5b)
# ---------------------------------------------------------------------------------------------------------------------
def _gen_rows(torch, make_pair, cfg, band, row_a, row_b, device):
    """
    Proc rows [row_a, row_b) of band `band` of the sharded raster: (source rows [(row_b - row_a) * ratio, W], reference
    rows [row_b - row_a, wp]).  Generated in global chunks of ROW_BAND_CHUNK proc rows seeded by (band, chunk), so every
    partition of the raster sees the same pixels (and no rank ever holds the 57.6 GB raster).
    """
    ratio, wp = cfg['ratio'], cfg['wp']
    tdtype = getattr(torch, cfg['dtype'])
    src = torch.empty(((row_b - row_a) * ratio, wp * ratio), dtype=tdtype, device=device)
    ref = torch.empty((row_b - row_a, wp), dtype=torch.float32, device=device)
    crs = tf = None
    for c in range(row_a // ROW_BAND_CHUNK, (max(row_b, row_a + 1) - 1) // ROW_BAND_CHUNK + 1):
        c0, c1 = c * ROW_BAND_CHUNK, min((c + 1) * ROW_BAND_CHUNK, cfg['hp'])
        s_ra, r_ra = make_pair(c1 - c0, wp, ratio, bands=1, dtype=cfg['dtype'], mu=cfg['mu'] * (1 + 0.15 * band),
                               seed=5000 + 131 * band + c, device=device, src_nodata=cfg['src_nodata'], ref_pad=0,
                               holes=2, bad_blobs=1)
        lo, hi = max(c0, row_a), min(c1, row_b)
        if hi > lo:
            src[(lo - row_a) * ratio:(hi - row_a) * ratio] = s_ra.array[0, (lo - c0) * ratio:(hi - c0) * ratio]
            ref[lo - row_a:hi - row_a] = r_ra.array[0, lo - c0:hi - c0]
        crs, tf = s_ra.crs, s_ra.transform
        del s_ra, r_ra
    return src, ref, crs, tf


class RowBandJob:
    """ The rows [a, b) of the sharded raster on this rank, and one step of the sharded path over all bands. """

    def __init__(self, torch, cfg, bands, rank, device, group):
        from homonim_b200 import Affine, KernelModel, Model, RasterArray, RefSpaceModel
        from homonim_b200 import dist as hd
        from homonim_b200.synthetic import make_pair
        self.torch, self.cfg, self.bands, self.rank, self.device, self.group, self.hd = torch, cfg, bands, rank, device, group, hd
        self.RasterArray, self.Affine = RasterArray, Affine
        a, b = bands.band(rank)
        self.a, self.b = a, b
        ratio, n_bands = cfg['ratio'], cfg['bands']
        self.same_grid = cfg['proc_crs'] == 'src'
        kw = dict(r2_inpaint_thresh=cfg['r2_inpaint_thresh'])
        self.model = (KernelModel if self.same_grid else RefSpaceModel)(Model(cfg['model']), cfg['kernel_shape'], **kw)
        self.halo = hd.halo_rows(cfg['kernel_shape'], proc_crs_ref=not self.same_grid, inpaint=False)
        self.src, self.ref = [], []
        for band in range(n_bands):
            s, r, crs, tf0 = _gen_rows(torch, make_pair, cfg, band, a, b, device)
            if self.same_grid:
                # planes allocated WITH their halo rows; the neighbours' rows are received in place every step
                s_ext, top = hd.alloc_with_halo(bands, rank, self.halo, cfg['wp'], torch.float32, device)
                r_ext, _ = hd.alloc_with_halo(bands, rank, self.halo, cfg['wp'], torch.float32, device)
                s_ext[top:top + (b - a)] = s
                r_ext[top:top + (b - a)] = r
                self.top = top
                s, r = s_ext, r_ext
            self.src.append(s)
            self.ref.append(r)
            torch.cuda.empty_cache()
        self.crs = crs
        res, x0, y0 = tf0.a, tf0.c, tf0.f
        self.ref_global_tf = Affine(res * ratio, 0.0, x0, 0.0, -res * ratio, y0)
        self.src_local_tf = Affine(res, 0.0, x0, 0.0, -res, y0 - a * ratio * res)
        if not self.same_grid:
            # the (small) reference planes are replicated: gather every rank's rows once, outside the timed region
            self.ref = [hd.all_gather_rows(r, bands, group) if bands.starts[-1] != (b - a) else r for r in self.ref]
        self.local_px = (b - a) * ratio * cfg['wp'] * ratio
        n_streams = int(os.environ.get('HB_ROW_BAND_STREAMS', '4'))
        self.streams = [torch.cuda.Stream(device=device) for _ in range(max(1, min(n_streams, cfg['bands'])))]
        # the small kernels / exchanges between a band's two streaming kernels run on high-priority twins of the streams
        self.hi_streams = [torch.cuda.Stream(device=device, priority=-1) for _ in self.streams]
        self.out = [torch.empty(((b - a) * ratio, cfg['wp'] * ratio), dtype=torch.float32, device=device)
                    for _ in range(len(self.streams))]
        self.stagger = os.environ.get('HB_ROW_BAND_STAGGER', '1') == '1'

    def step(self, keep_band0=None, serial=False):
        """ All bands, each on its own stream (and output plane), so that one band's small kernels and exchanges overlap the
        others' streaming kernels (``serial``: one stream, for per-kernel timing).  ``keep_band0``: tensor receiving band
        0's result. """
        torch, cfg = self.torch, self.cfg
        from homonim_b200.kernel_model import current_stream, on_stream
        main = current_stream()
        for st in self.streams + self.hi_streams:
            st.wait_stream(main)
        ns = len(self.streams)
        pick = (lambda band: 0) if serial else (lambda band: band % ns)
        outs = [keep_band0 if (band == 0 and keep_band0 is not None) else self.out[pick(band)]
                for band in range(cfg['bands'])]
        if self.same_grid:
            for band in range(cfg['bands']):
                with on_stream(self.streams[pick(band)]):
                    self.hd.fit_apply_same_grid_sharded(self.model, self.src[band], cfg['src_nodata'], self.ref[band],
                                                        NAN, self.bands, self.group, out=outs[band])
        else:
            # stage 1 of every band first (the down-sampling: a streaming kernel, no communication), then the stages 2:
            # while the host issues one band's statistics / exchanges / fit, the GPU still has the other bands' streaming
            # kernels queued.  (serial: band after band on one stream, for the per-kernel timing.)
            shards = {}

            def begin(band):
                with on_stream(self.streams[pick(band)]):
                    src_local = self.RasterArray(self.src[band], self.crs, self.src_local_tf, nodata=cfg['src_nodata'])
                    ref_ra = self.RasterArray(self.ref[band], self.crs, self.ref_global_tf, nodata=NAN)
                    shards[band] = self.hd.fuse_refspace_sharded_begin(self.model, src_local, ref_ra, self.bands,
                                                                       self.group)

            def end(band):
                if serial or not self.stagger:
                    with on_stream(self.streams[pick(band)]):
                        self.hd.fuse_refspace_sharded_end(shards.pop(band), out=outs[band])
                else:
                    with on_stream(self.hi_streams[pick(band)]):
                        self.hd.fuse_refspace_sharded_end(shards.pop(band), out=outs[band],
                                                          apply_stream=self.streams[pick(band)])

            if serial or not self.stagger:
                for band in range(cfg['bands']):
                    begin(band)
                    end(band)
            else:
                for band in range(cfg['bands']):
                    begin(band)
                for band in range(cfg['bands']):
                    end(band)
        for st in self.streams + self.hi_streams:
            main.wait_stream(st)


def measure_row_band(args, cfg, rank, world, local_rank, steps, warmup, with_n1=True):
    """ The one-raster regime on the current process group: returns the `row_band` object (rank 0) or None. """
    import torch
    import torch.distributed as dist
    from homonim_b200 import _native
    from homonim_b200.dist import RowBands
    from homonim_b200.kernel_model import KernelTimer
    device = torch.device('cuda', local_rank)
    lib = _native.lib()
    group = dist.group.WORLD if world > 1 else None
    bands = RowBands.split(cfg['hp'], world)
    if world == 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29512')
        dist.init_process_group('nccl', device_id=device, rank=0, world_size=1)
    solo = dist.new_group([0]) if world > 1 else None            # (collective: every rank takes part in its creation)
    t_gen0 = time.perf_counter()
    job = RowBandJob(torch, cfg, bands, rank, device, group)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t_gen0
    npix_total = cfg['hp'] * cfg['wp'] * cfg['ratio'] ** 2 * cfg['bands']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(j, n_steps, n_warm, sync):
        for _ in range(n_warm):
            j.step()
        sync()
        lib.hb_reset_launch_count()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        t_host = time.perf_counter()
        start.record()
        for _ in range(n_steps):
            j.step()
        end.record()
        host_ms[0] = (time.perf_counter() - t_host) * 1e3 / n_steps      # host time to ENQUEUE a step (no sync inside)
        sync()
        return start.elapsed_time(end), lib.hb_launch_count()

    elapsed_ms, launches = timed(job, steps, warmup, barrier)
    host_enqueue_ms = host_ms[0]
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = npix_total * steps / (elapsed_ms * 1e-3) / 1e6
    # per-kernel durations: one more step with CUDA events around every native call (bands then run back to back)
    with KernelTimer() as timer:
        job.step(serial=True)
        kernel_ms = timer.results()
    if os.environ.get('HB_ROW_BAND_TRACE') == '1':
        # where the native calls of one concurrent step ran on the device's clock (diagnostics; EVERY rank runs the step)
        barrier()
        barrier_ev = torch.cuda.Event(enable_timing=True)
        with KernelTimer() as tracer:
            barrier_ev.record()
            job.step()
            lines = tracer.timeline(barrier_ev)
        if rank == 0:
            for name, t0_, t1_ in lines:
                print(f'trace {name:32s} {t0_:8.3f} -> {t1_:8.3f} ms', file=sys.stderr)
    b_in = 4
    # algorithmic bytes per launch: per SOURCE pixel of this rank's rows for the resampling kernels, per proc-grid pixel
    # for the statistics / same-grid kernels
    proc_px = (bands.size(rank) * cfg['wp'])
    alg = {'hb_upsample_apply': job.local_px * (b_in + 4), 'hb_downsample_average': job.local_px * b_in,
           'hb_fit_apply_same_grid_rows': proc_px * 12, 'hb_block_norm_partial': proc_px * 8}
    kernels = {}
    for k, v in kernel_ms.items():
        kernels[k] = {'launches': len(v), 'ms_avg': round(sum(v) / len(v), 4)}
        if k in alg:
            kernels[k]['gbs'] = round(alg[k] / (kernels[k]['ms_avg'] * 1e-3) / 1e9, 1)
    bytes_per_px = 12 if not job.same_grid else 12 + 3 * 8       # DESIGN.md section 4 (same grid: + 3 statistics passes)
    peak_gbs, _ = _peak()
    result = {
        'workload': cfg['desc'], 'value': round(value, 1), 'unit': 'Mpix/s', 'scaling': 'strong', 'n_gpus': world,
        'steps': steps, 'warmup': warmup, 'ms_per_step': round(elapsed_ms / steps, 4), 'gpu_launches': int(launches),
        'pixels_per_step': int(npix_total), 'generate_s': round(gen_s, 1),
        'host_enqueue_ms_per_step': round(host_enqueue_ms, 3), 'streams': len(job.streams), 'staggered': bool(job.stagger),
        'sharding': f'row bands of {cfg["hp"]} proc rows over {world} rank(s); per band: 3 all-gathers of 131 KB of '
                    f'block statistics + P2P halo rows ({job.halo} proc rows per side); no rank reads another rank\'s '
                    f'pixels',
        'whole_step': {'algorithmic_bytes_per_pixel': bytes_per_px,
                       'achieved_gbs_per_gpu': round(value * 1e6 * bytes_per_px / 1e9 / world, 1),
                       'frac_of_measured_peak': round(value * 1e6 * bytes_per_px / 1e9 / world / peak_gbs, 4)},
        'kernels': kernels,
    }
    # ---- N > 1: the same raster on ONE GPU (rank 0, the other ranks wait), efficiency and stripe parity ----------------
    if world > 1 and with_n1 and not job.same_grid:      # (same grid: two full-size float32 planes per band -- the
        a, b = bands.band(rank)                          #  unsharded raster does not fit one GPU beside the outputs)
        rows_src = (b - a) * cfg['ratio']
        stripe = min(2048, rows_src)
        shard0 = torch.empty((rows_src, cfg['wp'] * cfg['ratio']), dtype=torch.float32, device=device)
        job.step(keep_band0=shard0)                              # band 0 of the sharded result
        torch.cuda.synchronize()
        cuts = sorted({0, (world - 1) // 2, world - 2})          # boundaries g | g + 1
        del job
        torch.cuda.empty_cache()
        n1 = parity = None
        if rank == 0:
            job1 = RowBandJob(torch, cfg, RowBands.split(cfg['hp'], 1), 0, device, solo)
            n1_steps = max(2, min(steps, 3))
            ms1, _ = timed(job1, n1_steps, 2, torch.cuda.synchronize)
            full0 = torch.empty((cfg['hp'] * cfg['ratio'], cfg['wp'] * cfg['ratio']), dtype=torch.float32, device=device)
            job1.step(keep_band0=full0)
            torch.cuda.synchronize()
            v1 = npix_total * n1_steps / (ms1 * 1e-3) / 1e6
            n1 = {'value': round(v1, 1), 'ms_per_step': round(ms1 / n1_steps, 4), 'steps': n1_steps,
                  'note': 'the same raster unsharded on rank 0 (a one-rank group), measured in this process'}
            del job1
        # stripes straddling shard boundaries, sent to rank 0
        worst, same_px, tot_px, masks_ok = 0.0, 0, 0, True
        for g in cuts:
            for side, owner in ((0, g), (1, g + 1)):
                oa, ob = bands.band(owner)
                n_rows = min(stripe, (ob - oa) * cfg['ratio'])
                if n_rows == 0:
                    continue
                row0 = oa * cfg['ratio'] + ((ob - oa) * cfg['ratio'] - n_rows if side == 0 else 0)
                if rank == owner:
                    piece = shard0[(rows_src - n_rows):] if side == 0 else shard0[:n_rows]
                    if owner != 0:
                        dist.send(piece.contiguous(), 0)
                if rank == 0:
                    if owner == 0:
                        got = piece
                    else:
                        got = torch.empty((n_rows, cfg['wp'] * cfg['ratio']), dtype=torch.float32, device=device)
                        dist.recv(got, owner)
                    exp = full0[row0:row0 + n_rows]
                    masks_ok = masks_ok and bool(torch.equal(torch.isnan(got), torch.isnan(exp)))
                    fin = torch.isfinite(exp)
                    floor = 1e-3 * exp[fin].abs().mean()
                    rel = ((got - exp).abs()[fin] / exp.abs()[fin].clamp_min(floor)).max().item()
                    worst = max(worst, rel)
                    same_px += int(((got == exp) | (torch.isnan(got) & torch.isnan(exp))).sum().item())
                    tot_px += got.numel()
        if rank == 0:
            parity = {'against': 'the unsharded single-GPU result of the same raster (band 1)',
                      'stripes': f'{len(cuts)} stripes of 2 x {stripe} source rows straddling the boundaries after '
                                 f'rank(s) {cuts}', 'masks_identical': masks_ok,
                      'max_rel_err_corr': float(f'{worst:.3e}'),
                      'bit_identical_frac': round(same_px / max(tot_px, 1), 6), 'pixels': tot_px}
            result['n1'] = n1
            result['efficiency_vs_n1'] = round(value / (world * n1['value']), 4)
            result['parity'] = parity
        barrier()
    return result if rank == 0 else None


# ---------------------------------------------------------------------------------------------------------------------
# main
# ---------------------------------------------------------------------------------------------------------------------
def _bare_copy_bandwidth(torch, src_host, out_host, device):
    """ GB/s of the pinned host -> device and device -> host copies alone (what bounds `e2e`). """
    d_in = torch.empty(src_host.shape, dtype=src_host.dtype, device=device)
    d_out = torch.empty(out_host.shape, dtype=out_host.dtype, device=device)
    res = {}
    for name, fn, nbytes in (('h2d', lambda: d_in.copy_(src_host, non_blocking=True), src_host.numel() * src_host.element_size()),
                             ('d2h', lambda: out_host.copy_(d_out, non_blocking=True), out_host.numel() * out_host.element_size())):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        res[name + '_gbs'] = round(2 * nbytes / (time.perf_counter() - t0) / 1e9, 2)
    # both directions at once (two copy engines)
    s2 = torch.cuda.Stream(device=device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d_in.copy_(src_host, non_blocking=True)
    with torch.cuda.stream(s2):
        out_host.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res['both_gbs'] = round((src_host.numel() * src_host.element_size() + out_host.numel() * out_host.element_size()) / dt / 1e9, 2)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-row-band', action='store_true', help='skip the one-raster (C5a) measurement')
    ap.add_argument('--row-band-workload', default='c5a', choices=['c5a', 'c5b', 'tiny-band'])
    ap.add_argument('--no-row-band-n1', action='store_true',
                    help='N > 1: skip the single-GPU run of the same raster on rank 0 (efficiency / stripe parity)')
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
    else:
        # every rank keeps to its own share of the host cores (the ranks' launch threads otherwise migrate over each other)
        try:
            cores = sorted(os.sched_getaffinity(0))
            share = max(1, len(cores) // world)
            os.sched_setaffinity(0, cores[local_rank * share:(local_rank + 1) * share] or cores)
        except (AttributeError, OSError):
            pass

    import torch
    import torch.distributed as dist
    from homonim_b200 import Model, ProcCrs, RasterArray, RasterFuse, _native
    from homonim_b200.kernel_model import KernelTimer
    from homonim_b200.synthetic import make_pair

    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1 or cfg.get('sharded'):
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29511')
        opts = None
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            opts = None
        dist.init_process_group('nccl', device_id=device, rank=rank, world_size=world, pg_options=opts)
    lib = _native.lib()

    if cfg.get('sharded'):
        # the one-raster regime as the headline
        stop, samples = threading.Event(), []
        sampler = threading.Thread(target=_clock_sampler, args=(stop, samples, local_rank, _gpu_uuid(torch, local_rank)),
                                   daemon=True)
        if rank == 0:
            sampler.start()
        t0 = time.perf_counter()
        rb = measure_row_band(args, cfg, rank, world, local_rank, args.steps, args.warmup, with_n1=not args.no_row_band_n1)
        t1 = time.perf_counter()
        stop.set()
        if rank == 0:
            sampler.join(timeout=2)
            peak_gbs, peak_src = _peak()
            dom = max((k for k in rb['kernels'] if 'gbs' in rb['kernels'][k]),
                      key=lambda k: rb['kernels'][k]['ms_avg'] * rb['kernels'][k]['launches'])
            line = {
                'metric': 'fit+apply Mpix/s', 'value': rb['value'], 'unit': 'Mpix/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': rb['ms_per_step'],
                'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': _config(cfg), 'clocks': _clocks_summary(samples, (t0, t1)), 'e2e': None,
                'gpu_launches': rb['gpu_launches'],
                'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': rb['kernels'][dom]['gbs'], 'peak': peak_gbs,
                             'peak_source': peak_src, 'unit': 'GB/s',
                             'frac': round(rb['kernels'][dom]['gbs'] / peak_gbs, 4), 'traffic': None,
                             'kernels': rb['kernels'], 'whole_step': rb['whole_step']},
                'cpu_baseline': None, 'row_band': rb,
            }
            print(json.dumps(line), flush=True)
        dist.destroy_process_group()
        return

    # ---- inputs: every rank corrects its own source image (seeded per rank), resident in HBM ------------------------
    if cfg.get('fixture'):
        src_ra, ref_ra = _fixture_pair(cfg, device)
    else:
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
        # the timed region lasts a few milliseconds: wait until the sampler is up (NVML initialisation takes longer than
        # that the first time) and keep the GPU under load meanwhile, so that its first samples are already "under load"
        t_wait = time.perf_counter()
        while not samples and time.perf_counter() - t_wait < 5.0:
            step()
            torch.cuda.synchronize()
    lib.hb_reset_launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_region0 = time.perf_counter()
    if _small(cfg):
        # inputs fit in L2: flush it between steps (a 512 MB fill) and time every step on its own
        flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=device)
        pairs = []
        for i in range(args.steps):
            flush_buf.fill_(i & 0xff)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            b.record()
            pairs.append((a, b))
        barrier()
        elapsed_ms = sum(a.elapsed_time(b) for a, b in pairs)
        del flush_buf
    else:
        start.record()
        for _ in range(args.steps):
            step()
        end.record()
        barrier()
        elapsed_ms = start.elapsed_time(end)
    launches = lib.hb_launch_count()
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

        def e2e_run(out_dtype, out_nodata):
            out_host = torch.empty(src_ra.array.shape, dtype=getattr(torch, out_dtype), pin_memory=True)
            out_profile = dict(dtype=out_dtype, nodata=out_nodata)

            def e2e_step():
                # the public call with HOST rasters (pinned CPU tensors): per band, host -> device copy of the source,
                # fit + apply, device -> host copy of the corrected band into `corr_out`; process() returns when the
                # corrected image is in host memory
                s = RasterArray(src_host, src_ra.crs, src_ra.transform, nodata=src_ra.nodata)
                r = RasterArray(ref_host, ref_ra.crs, ref_ra.transform, nodata=ref_ra.nodata)
                with RasterFuse(s, r, proc_crs=ProcCrs(cfg['proc_crs'])) as f:
                    f.process(model=Model(cfg['model']), kernel_shape=cfg['kernel_shape'], model_config=model_config,
                              corr_out=out_host, out_profile=out_profile)

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
            h2d = src_host.numel() * src_host.element_size() + ref_host.numel() * ref_host.element_size()
            d2h = out_host.numel() * out_host.element_size()
            res = {'value': round(npix * world * e_steps / (float(te.item()) * 1e-3) / 1e6, 2), 'unit': 'Mpix/s',
                   'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h), 'steps': e_steps,
                   'out_dtype': out_dtype}
            bw = _bare_copy_bandwidth(torch, src_host, out_host, device)
            tb = torch.tensor([bw['h2d_gbs'], bw['d2h_gbs'], bw['both_gbs']], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(tb, op=dist.ReduceOp.MIN)
            res['bare_copy_gbs_min_over_ranks'] = {'h2d': round(float(tb[0]), 2), 'd2h': round(float(tb[1]), 2),
                                                   'both_directions': round(float(tb[2]), 2)}
            # what the copies alone allow: the step cannot be faster than moving its bytes over PCIe
            res['copy_bound_mpix_s'] = round(npix * world / ((h2d + d2h) / (float(tb[2]) * 1e9)) / 1e6, 1)
            del out_host
            return res

        e2e = e2e_run('float32', NAN)                          # the reference's default output profile
        if cfg['dtype'] in ('uint8', 'uint16'):
            # the output in the source's own dtype (out_profile dtype / nodata, fuse.py:114-149): a quarter / half of
            # the device -> host bytes, conversion fused into the apply kernel's epilogue
            e2e['same_dtype_output'] = e2e_run(cfg['dtype'], 0)
        e2e['limiter'] = 'host <-> device copies (PCIe): compare `value` with `copy_bound_mpix_s`'
        del src_host, ref_host

    # ---- roofline of the dominant kernel (per-launch CUDA-event durations from the timed region) ---------------------
    peak_gbs, peak_src = _peak()
    b_in = src_ra.array.element_size()
    band_px = npix // cfg['bands']
    proc_px = (cfg['hp'] * cfg['wp']) if cfg['proc_crs'] == 'ref' else band_px
    want_r2 = cfg['model'] == 'gain-offset' and cfg['r2_inpaint_thresh'] is not None
    alg_bytes = {                                    # algorithmic bytes per launch
        'hb_upsample_apply': band_px * (b_in + 4) + 8 * proc_px,
        'hb_downsample_average': band_px * b_in + 4 * proc_px,
        'hb_fit_same_grid_rows': proc_px * (8 + 4 * (3 if want_r2 else 2)),
        'hb_fit_apply_same_grid': proc_px * 12,
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
                'traffic_source': 'static: profiles/ncu_traffic.json (ncu --set full capture committed with the repo)',
                'algorithmic_bytes_per_launch': int(alg_bytes[dominant]),
                'share_of_step': round(sum(kernel_ms[dominant]) / serial_ms, 3),
                'whole_step': {'algorithmic_bytes_per_pixel': step_bytes_per_px,
                               'achieved_gbs': round(value * 1e6 * step_bytes_per_px / 1e9 / world, 1),
                               'frac': round(value * 1e6 * step_bytes_per_px / 1e9 / world / peak_gbs, 4),
                               'frac_of_nominal_8000_gbs': round(value * 1e6 * step_bytes_per_px / 1e9 / world / 8000.0, 4),
                               'note': 'per GPU: value x algorithmic bytes per source band-pixel of the whole fit + '
                                       'apply step, against the same peak'},
                'timing': 'CUDA events around every launch, K steps with the bands serialised on one stream '
                          f'({round(serial_ms / args.steps, 4)} ms/step); `value` runs the bands on concurrent streams',
                'kernels': per_kernel}

    # ---- CPU baseline + parity: the reference on a bounded sample (rank 0, N = 1 only) ---------------------------------
    cpu_baseline = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, parity = measure_parity_and_cpu(args, cfg, src_ra, ref_ra)

    # ---- the one-raster regime in the same invocation ---------------------------------------------------------------------
    row_band = None
    if not args.no_row_band and args.workload in ('c2', 'tiny'):
        rb_cfg = WORKLOADS['tiny-band' if args.workload == 'tiny' else args.row_band_workload]
        fuse.close()
        del fuse, src_ra, ref_ra
        torch.cuda.empty_cache()
        row_band = measure_row_band(args, rb_cfg, rank, world, local_rank, steps=max(3, min(args.steps, 5)), warmup=3,
                                    with_n1=not args.no_row_band_n1)

    if rank == 0:
        line = {
            'metric': 'fit+apply Mpix/s', 'value': round(value, 1), 'unit': 'Mpix/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(elapsed_ms / args.steps, 4),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': _data(cfg),
            'config': _config(cfg), 'sharding': 'one source image per GPU, no collective',
            'clocks': _clocks_summary(samples, (t_region0, t_region1)), 'e2e': e2e, 'gpu_launches': int(launches),
            'roofline': roofline, 'parity': parity, 'cpu_baseline': cpu_baseline, 'row_band': row_band,
        }
        print(json.dumps(line), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

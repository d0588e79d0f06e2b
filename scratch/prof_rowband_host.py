"""
Host-side profile of one row-band step (N = 1, C5a geometry scaled by argv[1] proc rows): where the enqueue time goes.
usage: python scratch/prof_rowband_host.py [hp] [steps]
"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, '.')
import torch
import torch.distributed as dist

import bench
from homonim_b200.dist import RowBands

hp = int(sys.argv[1]) if len(sys.argv) > 1 else 750
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
os.environ.setdefault('MASTER_PORT', '29533')
device = torch.device('cuda', 0)
torch.cuda.set_device(0)
dist.init_process_group('nccl', device_id=device, rank=0, world_size=1)
cfg = dict(bench.WORKLOADS['c5a'], hp=hp)
bands = RowBands.split(cfg['hp'], 1)
job = bench.RowBandJob(torch, cfg, bands, 0, device, None)
for _ in range(3):
    job.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    job.step()
t_host = (time.perf_counter() - t0) / steps * 1e3
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / steps * 1e3
print(f'host enqueue {t_host:.3f} ms / step, wall {t_all:.3f} ms / step')
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    job.step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(45)
st.sort_stats('tottime').print_stats(25)
dist.destroy_process_group()

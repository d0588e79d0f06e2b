"""
File-to-file run of the C2 geometry: where the time goes when RasterFuse is given GeoTIFF file names.
usage: python scratch/perf_files.py [proc_px=250] [compress=deflate|none] [out_dtype=float32|uint16]
Writes a 4-band uint16 source of (20 * proc_px)^2 pixels and its reference to /tmp, corrects it file -> file, and prints the
wall time of the three stages: read + decode (FilePair), GPU step incl. host <-> device copies, encode + write.
"""
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, '.')
import homonim_b200.files as hf
import homonim_b200.fuse as hfuse
from homonim_b200 import Model, RasterFuse
from homonim_b200.geotiff import write_geotiff
from homonim_b200.synthetic import make_pair

proc = int(sys.argv[1]) if len(sys.argv) > 1 else 250
compress = sys.argv[2] if len(sys.argv) > 2 else 'deflate'
out_dtype = sys.argv[3] if len(sys.argv) > 3 else 'float32'
src_ra, ref_ra = make_pair(proc, proc, 20, bands=4, dtype='uint16', mu=3000.0, seed=2, device='cuda', src_nodata=0.0)
tmp = tempfile.mkdtemp(prefix='hb_files_')
geokeys = ((1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, 32735), (), '')
wl = (0.48, 0.56, 0.66, 0.83)
t0 = time.perf_counter()
src_path = write_geotiff(os.path.join(tmp, 'src.tif'), src_ra.to_host().array, src_ra.transform, nodata=0, geokeys=geokeys,
                         band_tags=[dict(center_wavelength=w) for w in wl], compress=None if compress == 'none' else compress)
ref_path = write_geotiff(os.path.join(tmp, 'ref.tif'), ref_ra.to_host().array, ref_ra.transform, geokeys=geokeys,
                         band_tags=[dict(center_wavelength=w) for w in wl], compress=None if compress == 'none' else compress)
print(f'inputs written in {time.perf_counter() - t0:.1f} s: {os.path.getsize(src_path) / 1e6:.0f} MB source '
      f'({src_ra.array.numel() * 2 / 1e6:.0f} MB raw), cores {os.cpu_count()}')
npix = src_ra.array.numel()
del src_ra, ref_ra
torch.cuda.empty_cache()

stage = {}
orig_init = hf.FilePair.__init__
orig_write = hf.write_geotiff


def timed_init(self, *a, **k):
    t = time.perf_counter()
    orig_init(self, *a, **k)
    stage['read + decode'] = time.perf_counter() - t


def timed_write(*a, **k):
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = orig_write(*a, **k)
    stage['encode + write'] = stage.get('encode + write', 0.0) + time.perf_counter() - t
    return r


hf.FilePair.__init__ = timed_init
hf.write_geotiff = timed_write
for rep in range(2):
    stage.clear()
    out = os.path.join(tmp, f'corr{rep}.tif')
    t0 = time.perf_counter()
    with RasterFuse(src_path, ref_path) as fuse:
        fuse.process(out, Model.gain_offset, (15, 15), build_ovw=False,
                     out_profile=dict(dtype=out_dtype, nodata=(0 if out_dtype != 'float32' else float('nan')),
                                      creation_options=dict(tiled=True, blockxsize=512, blockysize=512, compress=compress,
                                                            interleave='band', photometric=None)))
    total = time.perf_counter() - t0
    gpu = total - sum(stage.values())
    print(f'run {rep}: total {total:.2f} s = {npix / total / 1e6:.1f} Mpix/s | read + decode {stage["read + decode"]:.2f} s | '
          f'GPU step incl. H2D / D2H {gpu:.2f} s | encode + write {stage["encode + write"]:.2f} s '
          f'({os.path.getsize(out) / 1e6:.0f} MB {out_dtype} {compress})')

import torch, sys, os
sys.path.insert(0, '.')
from homonim_b200 import kernel_model as hkm
from homonim_b200.geometry import Affine
def timeit(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    ev=[torch.cuda.Event(enable_timing=True) for _ in range(n+1)]
    ev[0].record()
    for i in range(n):
        f(); ev[i+1].record()
    torch.cuda.synchronize()
    ts=sorted(ev[i].elapsed_time(ev[i+1]) for i in range(n))
    return ts[n//2]
for n, dt in ((10000, 'u16'), (10000, 'f32'), (10000, 'u8')):
    src = torch.randint(1, 250, (n, n), device='cuda', dtype=torch.int32)
    src = src.to({'u16': torch.uint16, 'f32': torch.float32, 'u8': torch.uint8}[dt])
    src_tf = Affine(0.5, 0, 0, 0, -0.5, 0); ref_tf = Affine(10, 0, 0, 0, -10, 0)
    nd = float('nan') if dt == 'f32' else 0
    t = timeit(lambda: hkm._downsample_average(src, src_tf, nd, (n//20, n//20), ref_tf))
    print(os.environ.get('HOMONIM_B200_LIB', 'default')[-16:], f'downsample {dt} {n}: {t*1e3:.1f} us  {src.numel()*src.element_size()/t/1e9:.2f} TB/s')

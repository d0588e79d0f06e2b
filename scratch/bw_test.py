import torch, sys
sys.path.insert(0, '.')
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    ev=[torch.cuda.Event(enable_timing=True) for _ in range(n+1)]
    ev[0].record()
    for i in range(n):
        f(); ev[i+1].record()
    torch.cuda.synchronize()
    ts=sorted(ev[i].elapsed_time(ev[i+1]) for i in range(n))
    return ts[n//2]
for mb in (200, 800):
    n = mb * 1000 * 1000 // 4
    x = torch.rand(n, device='cuda'); y = torch.empty_like(x)
    t = timeit(lambda: x.sum()); print(f'{mb} MB  sum (read only): {t*1e3:.1f} us  {mb/t/1e3*1e3/1e3:.2f} TB/s')
    t = timeit(lambda: y.copy_(x)); print(f'{mb} MB  copy: {t*1e3:.1f} us  {2*mb/t/1e3:.2f} TB/s')
    t = timeit(lambda: y.fill_(1.0)); print(f'{mb} MB  fill (write only): {t*1e3:.1f} us  {mb/t/1e3:.2f} TB/s')
    xi = (x * 1000).to(torch.int16)
    t = timeit(lambda: xi.sum()); print(f'{mb//2} MB int16 sum: {t*1e3:.1f} us  {mb/2/t/1e3:.2f} TB/s')
# our downsample at two sizes
from homonim_b200 import kernel_model as hkm
from homonim_b200.geometry import Affine
for n in (10000, 20000):
    src = torch.randint(1, 5000, (n, n), device='cuda', dtype=torch.int32).to(torch.uint16)
    src_tf = Affine(0.5, 0, 0, 0, -0.5, 0); ref_tf = Affine(10, 0, 0, 0, -10, 0)
    t = timeit(lambda: hkm._downsample_average(src, src_tf, 0, (n//20, n//20), ref_tf))
    print(f'downsample u16 {n}: {t*1e3:.1f} us  {n*n*2/t/1e9:.2f} TB/s')
    srcf = src.to(torch.float32)
    t = timeit(lambda: hkm._downsample_average(srcf, src_tf, float("nan"), (n//20, n//20), ref_tf))
    print(f'downsample f32 {n}: {t*1e3:.1f} us  {n*n*4/t/1e9:.2f} TB/s')
    del src, srcf

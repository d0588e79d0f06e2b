import sys, torch
sys.path.insert(0, '.')
from homonim_b200 import KernelModel, Model
nan = float('nan')
km = KernelModel(Model.gain_blk_offset, (5, 5))
g = torch.Generator(device='cuda').manual_seed(0)
for (h, w) in ((400, 400), (711, 403), (1448, 1448)):
    src = torch.rand((h, w), generator=g, device='cuda') * 0.5 + 0.2
    ref = 0.7 * src + 0.05 + 0.01 * torch.rand((h, w), generator=g, device='cuda')
    src[10:30, 50:90] = nan
    for _ in range(3):
        norm = km._block_norm(src, nan, ref, nan)
    torch.cuda.synchronize()

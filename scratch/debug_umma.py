import sys; sys.path.insert(0, '.')
import torch
from dcase2019_task4_b200 import _lib
dev = torch.device('cuda', 0)
g = torch.Generator().manual_seed(1)
A = torch.randn(128, 64, generator=g).to(dev)
B = torch.randn(128, 64, generator=g).to(dev)
D = torch.full((128, 64), float('nan'), device=dev)
_lib.check(_lib.lib().dcase_selftest_umma(_lib.ctx(dev), 1, _lib.ptr(A), _lib.ptr(B), _lib.ptr(D), _lib.stream_ptr()))
torch.cuda.synchronize()
ref = A.double().t() @ B.double()   # [m][n]
raw = D.double()
print("nan lanes:", torch.isnan(raw).any(1).nonzero().flatten().tolist()[:20])
for l in list(range(0, 20)) + [32, 33, 48, 64, 96]:
    d = ((raw[l][None] - ref) ** 2).sum(1)
    dT = ((raw[l][None] - ref.t()) ** 2).sum(1)
    print(l, "best row", int(d.argmin()), float(d.min()), "| best col(transposed)", int(dT.argmin()), float(dT.min()), "| norm", float((raw[l]**2).sum()))
# partial sums hypotheses: only first 8 pixels etc
for npix in (8, 16, 32, 64, 128):
    r2 = A[:npix].double().t() @ B[:npix].double()
    print("npix", npix, float((raw[:16] - r2[:16]).abs().max()))

import sys, ctypes; sys.path.insert(0, '.')
import torch
from dcase2019_task4_b200 import _lib
dev = torch.device('cuda', 0)
L = _lib.lib()
L.dcase_selftest_umma_shift.restype = ctypes.c_int
L.dcase_selftest_umma_shift.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
g = torch.Generator().manual_seed(1)
A = torch.randn(256, 64, generator=g).to(dev)
B = torch.randn(64, 64, generator=g).to(dev)
for pitch in (8, 10, 16):
  for shift in (0, 1, 3, 4, 8, 11):
    if shift + 15 * pitch + 8 > 256: continue
    rows = torch.tensor([shift + (m // 8) * pitch + (m % 8) for m in range(128)], device=dev)
    ref = A[rows].double() @ B.double().t()
    out = []
    for mode in (0, 1):
        D = torch.zeros(128, 64, device=dev)
        _lib.check(L.dcase_selftest_umma_shift(_lib.ctx(dev), shift, pitch, mode, _lib.ptr(A), _lib.ptr(B), _lib.ptr(D), _lib.stream_ptr()))
        torch.cuda.synchronize()
        out.append(float((D.double() - ref).abs().max()))
    print(f"pitch {pitch} shift {shift}: err base_offset=0: {out[0]:.3e}   base_offset=(addr>>7)&7: {out[1]:.3e}   (max ref {float(ref.abs().max()):.1f})")

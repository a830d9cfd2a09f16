"""Pins the hand-written tcgen05 primitives (csrc/tc.cuh: SW128 operand layout, smem / instruction descriptors,
TMEM read-back) against a plain fp32 matmul.  kind::tf32 keeps 10 mantissa bits of each operand."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tf32(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)      # truncation to 19 bits, as the tensor core reads it


def _run(mode, A, B):
    from dcase2019_task4_b200 import _lib
    D = torch.full((128, 64), float("nan"), device=A.device)
    _lib.check(_lib.lib().dcase_selftest_umma(_lib.ctx(A.device), mode, _lib.ptr(A), _lib.ptr(B), _lib.ptr(D),
                                              _lib.stream_ptr()))
    torch.cuda.synchronize()
    return D


def test_umma_kmajor_128x64x64(cuda_device):
    g = torch.Generator().manual_seed(0)
    A = torch.randn(128, 64, generator=g).to(cuda_device)
    B = torch.randn(64, 64, generator=g).to(cuda_device)
    D = _run(0, A, B)
    ref = _tf32(A).double() @ _tf32(B).double().t()
    err = float((D.double() - ref).abs().max())
    full = float((D.double() - A.double() @ B.double().t()).abs().max())
    print(f"K-major tf32: err vs tf32-truncated ref {err:.3e}, vs fp32 ref {full:.3e}")
    assert err <= 2e-3 * float(ref.abs().max())                     # rounding mode of the operand conversion
    assert full <= 2e-2 * float(ref.abs().max())


def test_umma_mnmajor_m64(cuda_device):
    g = torch.Generator().manual_seed(1)
    A = torch.randn(128, 64, generator=g).to(cuda_device)
    B = torch.randn(128, 64, generator=g).to(cuda_device)
    raw = _run(1, A, B)                                             # [128 lanes][64 columns]
    ref = (_tf32(A).double().t() @ _tf32(B).double())               # [64 m][64 n]
    # M = 64 accumulators: row m lives in TMEM lane 32*(m/16) + m%16 (16 lanes of each 32-lane sub-partition)
    lanes = torch.tensor([32 * (m // 16) + m % 16 for m in range(64)], device=cuda_device)
    got = raw[lanes]
    err = float((got.double() - ref).abs().max())
    print(f"MN-major (SWIZZLE_128B_BASE32B) M=64: err with lane map 32*(m/16)+m%16: {err:.3e}")
    assert err <= 2e-3 * float(ref.abs().max())


@pytest.mark.parametrize("pitch,shift", [(8, 0), (8, 3), (10, 1), (10, 11), (16, 4)])
def test_umma_shifted_operand(cuda_device, pitch, shift):
    """The conv kernels address the 9 taps as 9 start addresses inside ONE staged halo: the operand swizzle must be
    a function of the absolute shared-memory address (descriptor base_offset = 0), for any row shift and any
    8-row-group pitch."""
    from dcase2019_task4_b200 import _lib
    g = torch.Generator().manual_seed(2)
    A = torch.randn(256, 64, generator=g).to(cuda_device)
    B = torch.randn(64, 64, generator=g).to(cuda_device)
    D = torch.zeros(128, 64, device=cuda_device)
    _lib.check(_lib.lib().dcase_selftest_umma_shift(_lib.ctx(cuda_device), shift, pitch, 0, _lib.ptr(A), _lib.ptr(B),
                                                    _lib.ptr(D), _lib.stream_ptr()))
    torch.cuda.synchronize()
    rows = torch.tensor([shift + (m // 8) * pitch + (m % 8) for m in range(128)], device=cuda_device)
    ref = _tf32(A[rows]).double() @ _tf32(B).double().t()
    assert float((D.double() - ref).abs().max()) <= 2e-3 * float(ref.abs().max())


def test_umma_issue_microbenchmark_matches_operand_floor(cuda_device):
    """dcase_bench_umma (tools/umma_bench.py): a tf32 MMA can never beat max(M,128) N / 256 cycles, nor the 128 B/cycle
    shared-memory operand fetch DESIGN.md section 3.1 builds on; the measured value stays within 25 % of that bound."""
    from dcase2019_task4_b200 import _lib
    lib, ctx = _lib.lib(), _lib.ctx(cuda_device)
    out = torch.zeros(4, device=cuda_device)
    stream = torch.cuda.current_stream().cuda_stream
    for M, N in ((128, 64), (128, 128), (64, 64)):
        assert lib.dcase_bench_umma(ctx, M, N, 0, 0, 1024, 4, 1024, 0, 0, out.data_ptr(), stream) == 0
        torch.cuda.synchronize()
        cyc = float(out.mean())
        bound = max(max(M, 128) * N / 256.0, (M + N) * 8 * 4 / 128.0)
        assert bound - 1.0 <= cyc <= 1.25 * bound, (M, N, cyc, bound)
    assert lib.dcase_bench_umma(ctx, 96, 64, 0, 0, 1024, 4, 1024, 0, 0, out.data_ptr(), stream) != 0     # illegal M


@pytest.mark.parametrize("mode", [2, 3])
def test_umma_a_operand_from_tensor_memory(cuda_device, mode):
    """A chained GEMM (A B^T) B^T whose second A operand is the first GEMM's accumulator, read in place from tensor memory
    (tc::umma_tf32_tmem_a_elect): an M = 128 accumulator already has the [lane = row][column = k] layout a TMEM A operand
    needs, so block 0's GLU linear can consume the conv output without a shared-memory round trip."""
    g = torch.Generator().manual_seed(5)
    A = torch.randn(128, 64, generator=g).to(cuda_device)
    B = (torch.randn(64, 64, generator=g) / 8).to(cuda_device)
    D = _run(mode, A, B)      # mode 3: the first accumulator is overwritten after the chained GEMM has completed
    y = (_tf32(A).double() @ _tf32(B).double().t()).float()
    ref = _tf32(y).double() @ _tf32(B).double().t()
    err = float((D.double() - ref).abs().max())
    print(f"TMEM-A chained GEMM: err vs tf32-truncated ref {err:.3e} (max |ref| {float(ref.abs().max()):.3e})")
    assert err <= 2e-3 * float(ref.abs().max())


def test_umma_bf16_mnmajor_m128_n16(cuda_device):
    """kind::f16 with bf16 operands, both MN-major in the 16-bit SWIZZLE_128B layout, M = 128 as two 64-wide blocks (LBO),
    N = 16, K = 128 in eight K = 16 instructions: the accumulating GEMM of cnn0's backward."""
    g = torch.Generator().manual_seed(5)
    A = torch.randn(128, 64, generator=g).to(cuda_device)
    B = torch.randn(128, 64, generator=g).to(cuda_device)
    raw = _run(4, A, B)[:, :16]                                     # [128 lanes = m][16 columns = j]
    Ab, Bb = A.bfloat16().double(), B.bfloat16().double()
    ref = torch.cat([Ab, Bb], 1).t() @ Ab[:, :16]                   # [128 m][16 j]
    err = float((raw.double() - ref).abs().max())
    print(f"bf16 MN-major M=128 N=16: err vs bf16-rounded ref {err:.3e} (max |ref| {float(ref.abs().max()):.3e})")
    assert err <= 1e-4 * float(ref.abs().max())                     # exact products, fp32 accumulation

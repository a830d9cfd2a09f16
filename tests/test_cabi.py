"""CPU: the C-ABI library loads and exports every symbol include/dcase_b200.h declares; shape / layout helpers
answer without a GPU; compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dcase_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcase_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from dcase2019_task4_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib


def test_every_declared_symbol_is_exported_and_bound(lib):
    handle = lib.lib()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(handle, name), f"{name} declared in include/dcase_b200.h but not exported"
        assert name in lib.SIGNATURES, f"{name} has no ctypes signature in _lib.SIGNATURES"
    assert set(lib.SIGNATURES) == set(syms)


def test_layout_helpers_without_gpu(lib):
    h = lib.lib()
    assert h.dcase_version() == 100
    assert h.dcase_logmel_num_frames(441000) == 864            # config.py:22 max_frames
    assert h.dcase_crnn_param_count(10) == 214356
    assert h.dcase_crnn_param_offset(10, b"cnn.cnn.conv0.weight") == 0
    assert h.dcase_crnn_param_offset(10, b"dense_softmax.bias") == 214356 - 10
    assert h.dcase_crnn_param_offset(10, b"no.such.param") == -1
    assert h.dcase_crnn_workspace_bytes(24, 864, 10) > 24 * 432 * 16 * 64 * 4 * 4
    off, n = ctypes.c_size_t(), ctypes.c_size_t()
    assert h.dcase_crnn_ws_tensor(24, 864, 10, b"out0", ctypes.byref(off), ctypes.byref(n)) == 0
    assert n.value == 24 * 432 * 16 * 64
    assert h.dcase_crnn_ws_tensor(24, 864, 10, b"bogus", ctypes.byref(off), ctypes.byref(n)) < 0
    assert b"bogus" in h.dcase_last_error()


def test_struct_mirrors_match_header(lib):
    h = lib.lib()
    assert ctypes.sizeof(lib.StepScalars) == h.dcase_sizeof_step_scalars() == 40      # 8 + 4 + 7 * 4
    assert ctypes.sizeof(lib.MtArgs) == h.dcase_sizeof_mt_args()


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.DcaseError):
        lib.ctx()
    h = ctypes.c_void_p()
    assert lib.lib().dcase_ctx_create(ctypes.byref(h), 0) < 0  # no device -> error code, not a silent CPU path
    from dcase2019_task4_b200.models.CRNN import CRNN
    from dcase2019_task4_b200 import config as cfg
    m = CRNN(**cfg.crnn_kwargs)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 864, 64))


def test_bigru_parameter_layout_matches_nn_gru():
    """Host arithmetic only: the experimental cluster BiGRU reads nn.GRU's tensors in named_parameters() order."""
    import torch
    from dcase2019_task4_b200 import _lib
    for H in (64, 128, 256):
        gru = torch.nn.GRU(64, H, num_layers=2, bidirectional=True, batch_first=True)
        assert _lib.lib().dcase_bigru_param_count_h(64, H) == sum(p.numel() for p in gru.parameters())
    assert _lib.lib().dcase_bigru_workspace_bytes_h(24, 108, 256) == 24 * 108 * (6 * 256 + 2 * 256) * 4
    assert _lib.lib().dcase_bigru_workspace_bytes(24, 108) == 24 * 108 * (6 * 64 + 2 * 64) * 4


def test_ctx_as_first_call_does_not_deadlock(monkeypatch):
    """Regression: ``_lib.ctx()`` used to take the module lock and then call ``lib()``, which takes it again -- a
    deadlock whenever a context was requested before anything else had loaded the library (the GPU Scaler tests run
    alone).  Here the library loads on the CPU, the context creation itself fails (no GPU) and must raise promptly."""
    import contextlib
    import threading
    import torch
    from dcase2019_task4_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_ctx", {})
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda idx: contextlib.nullcontext())
    outcome = {}

    def run():
        try:
            _lib.ctx()
            outcome["result"] = "created"
        except _lib.DcaseError as e:
            outcome["result"] = "raised: %s" % e
    t = threading.Thread(target=run, daemon=True)
    t.start()
    t.join(timeout=20)
    assert not t.is_alive(), "ctx() deadlocked on the library lock"
    if not torch.backends.cuda.is_built() or not _real_cuda_available():
        assert outcome["result"].startswith("raised")


def _real_cuda_available():
    import subprocess
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, timeout=10).returncode == 0
    except (OSError, subprocess.TimeoutExpired):
        return False


def test_fft_passes_of_the_logmel_kernel_on_the_host(tmp_path):
    """csrc/fft1024.cuh is __host__ __device__: the warp-autonomous real FFT the STFT kernel runs (two register-resident
    32-point DFTs, twiddle tree, 32 x 32 transpose indexing, Hermitian split) is compiled for the CPU and run lane by lane
    against a naive float64 DFT of a windowed 2048-sample frame, at even and odd offsets of the de-interleaved span
    (tests/csrc/fft_host_check.cu).  fp32 bound: 2e-6 of the largest bin."""
    import shutil
    import subprocess
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    exe = str(tmp_path / "fft_check")
    src = os.path.join(ROOT, "tests", "csrc", "fft_host_check.cu")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src], check=True,
                   capture_output=True, timeout=300)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    err, mag = float(r.stdout.split()[1]), float(r.stdout.split()[3])
    assert err <= 2e-6 * mag


def test_device_philox_code_on_the_host_matches_kat_and_numpy_restatement(tmp_path):
    """csrc/common.cuh::philox4x32_10 (the function the dropout / noise kernels call) compiled for the CPU: Random123
    known-answer vectors, and the same words as oracle/philox.py for the (row, stream, step, seed) counter layout of the
    RNG contract (include/dcase_b200.h)."""
    import shutil
    import subprocess
    import numpy as np
    from oracle import philox
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    exe = str(tmp_path / "philox_check")
    src = os.path.join(ROOT, "tests", "csrc", "philox_host_check.cu")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src], check=True,
                   capture_output=True, timeout=300)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    lines = r.stdout.strip().splitlines()
    assert r.returncode == 0 and lines[0] == "kat_failures 0", r.stdout + r.stderr
    seed = 0x0123456789abcdef
    for i, line in enumerate(lines[1:]):
        row, *words = (int(v) for v in line.split())
        want = philox.philox4x32(np.uint32(row & 0xFFFFFFFF), np.uint32(row >> 32), np.uint32(8 * 1 + 2), np.uint32(41 + i),
                                 np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32))
        assert [int(np.asarray(w).reshape(-1)[0]) for w in want] == words, (row, words)


def test_resampler_clock_equals_the_serial_float64_sum():
    """dcase_audio_resample's clock: resampy adds 1 / ratio to a float64 register once per output sample; the kernel
    rebuilds that serial sum from a per-binade table of linear segments (csrc/logmel.cu::build_resample_clock).  It must be
    the SAME doubles, bit for bit, or the samples whose exact time is an integer (every 147th at 48 -> 44.1 kHz) land on
    the other side of the algorithm's truncation jump."""
    import ctypes
    import numpy as np
    from dcase2019_task4_b200 import _lib
    L = _lib.lib()
    for sr_in, sr_out, n in ((48000, 44100, 441000), (44100, 16000, 160000), (16000, 44100, 441000), (22050, 44100, 99999),
                             (32000, 44100, 300001), (8000, 44100, 441000), (96000, 44100, 441000), (44100, 48000, 7)):
        out = np.empty(n)
        assert L.dcase_audio_resample_clock(n, sr_in, sr_out, ctypes.c_void_p(out.ctypes.data)) == 0
        inc = 1.0 / (float(sr_out) / float(sr_in))
        serial = np.concatenate([[0.0], np.cumsum(np.full(n - 1, inc))])      # numpy's cumsum is the serial sum
        assert np.array_equal(out, serial), (sr_in, sr_out)

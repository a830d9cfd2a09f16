"""Data parallelism on real GPUs: two ranks (one process per GPU) through ``MeanTeacherEngine`` -- VERDICT r1 weak 3.

After three iterations on different shards the replicas must be BIT-identical, and equal to a single-GPU
``dcase_adam_ema_step`` on the mean of the two ranks' gradients (sum in rank order, 1/N folded into the optimizer
kernel -- the same arithmetic the exchange performs).  Modes: NCCL all-reduce launched eagerly (DCASE_DP_NCCL=1) and the
default, the fused exchange + Adam + EMA kernel over NVLink peer memory (``csrc/p2p.cu``), eager and inside the step's
CUDA graph.  (Capturing the NCCL all-reduce into the graph hung on hardware in rounds 1 and 2: not offered.)
Needs two GPUs (``gpurun --gpus 2``); skipped on a one-GPU box.  The CPU side of the same logic: tests/test_dp_gloo.py.
"""
import os
import socket
import traceback

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        if not mode.startswith("p2p"):
            os.environ["DCASE_DP_NCCL"] = "1"
        import torch.distributed as dist
        import dcase2019_task4_b200.config as cfg
        from dcase2019_task4_b200 import kernels as K
        from dcase2019_task4_b200 import main as bmain
        from dcase2019_task4_b200 import dp, synth
        from dcase2019_task4_b200.models.CRNN import CRNN
        from oracle import crnn as ocrnn
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", device_id=dev)
        B, T, L, N = 8, 64, 511 * 64, 3
        waves, _ = synth.make_clips(N * B, seed=100 + rank, n_samples=L)          # a different shard per rank
        waves = torch.from_numpy(waves.reshape(N, B, L)).to(dev)
        tgt = (torch.rand(N, B, T // 8, 10, generator=torch.Generator().manual_seed(7 + rank)) < 0.2).float()
        tgt[:, 2:6] = -1
        tgt = tgt.to(dev)
        mean = torch.full((64,), -30.0, device=dev)
        std = torch.full((64,), 12.0, device=dev)
        ps, pt = ocrnn.init_params(seed=81), ocrnn.init_params(seed=82)
        student, teacher = CRNN(**cfg.crnn_kwargs), CRNN(**cfg.crnn_kwargs)
        with torch.no_grad():
            for m, p in ((student, ps), (teacher, pt)):
                for k, v in m.named_parameters():
                    v.copy_(p[k])
        for p_ in teacher.parameters():
            p_.detach_()
        student, teacher = student.train().cuda(), teacher.train().cuda()
        dp.broadcast_model_(student)
        dp.broadcast_model_(teacher)
        student._rng_seed, student._rng_step = 1000 + rank, 0
        opt = torch.optim.Adam(student.parameters(), lr=0.001, betas=(0.9, 0.999))
        eng = bmain.MeanTeacherEngine(student, opt, teacher, slice(2), slice(6, 8), B, T, use_graph=mode.endswith("graph"))
        n = student.flat_parameters().numel()
        # reference replica: the optimizer kernel alone, fed with the gathered per-rank gradients
        rp, re = student.flat_parameters().clone(), teacher.flat_parameters().clone()
        rm, rv = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        local = []
        if not mode.endswith("graph"):
            real = K.mt_fwd_bwd

            def spy(args):                       # the gradient slab right after the backward, before the exchange
                real(args)
                local.append(eng.grads.clone())
            bmain.K.mt_fwd_bwd = spy
        exact = 0.0
        for i in range(N):
            eng.step_from_waveforms(waves[i], tgt[i], mean, std, 0.5, i + 1, check=False)
            torch.cuda.synchronize()
            if not mode.endswith("graph"):
                parts = [torch.empty(n, device=dev) for _ in range(world)]
                dist.all_gather(parts, local[-1])
                total = parts[0].clone()
                for r in range(1, world):
                    total += parts[r]
                K.adam_ema_step(rp, total, rm, rv, re, i + 1, ema_alpha=min(1 - 1 / (i + 2), 0.999),
                                grad_scale=1.0 / world)
                torch.cuda.synchronize()
                exact = max(exact, float((rp - student.flat_parameters()).abs().max()),
                            float((re - teacher.flat_parameters()).abs().max()))
        both = [torch.empty(n, device=dev) for _ in range(world)]
        dist.all_gather(both, student.flat_parameters().detach())
        both_t = [torch.empty(n, device=dev) for _ in range(world)]
        dist.all_gather(both_t, teacher.flat_parameters().detach())
        identical = all(torch.equal(both[0], b) for b in both[1:]) and all(torch.equal(both_t[0], b) for b in both_t[1:])
        moved = float((both[0] - torch.cat([ps[k].reshape(-1) for k in ps]).to(dev)).abs().max())
        loss = eng.read_meters()["Loss"]
        q.put((rank, "ok", identical, exact, moved, loss, both[0].cpu() if rank == 0 else None))
        dist.barrier()
        if eng.p2p is not None:
            eng.p2p.close()
        dist.destroy_process_group()
    except Exception:
        q.put((rank, "error", traceback.format_exc()))


def _run(mode, world=2):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = sorted([q.get(timeout=150) for _ in range(world)], key=lambda r: r[0])
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()                      # a missed flag / stuck collective must not outlive the test
    for r in res:
        assert r[1] == "ok", r[2]
    return res


@pytest.fixture(scope="module")
def two_gpus(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")


@pytest.fixture(scope="module")
def eager_result(two_gpus):
    return _run("eager")


def test_two_nccl_ranks_stay_bit_identical_and_match_the_mean_gradient_update(eager_result):
    for rank, _, identical, exact, moved, loss, _ in eager_result:
        assert identical, "replicas diverged"
        assert exact == 0.0, "DP update differs from dcase_adam_ema_step on the summed gradients x 1/N"
        assert 1e-4 < moved <= 3.1e-3 and 0 < loss < 1e3


def test_two_ranks_with_the_fused_p2p_exchange_and_optimizer(eager_result):
    res = _run("p2p")
    for rank, _, identical, exact, moved, loss, flat in res:
        assert identical, "replicas diverged"
        assert exact <= 2.5e-7, "fused exchange + Adam + EMA differs from the optimizer kernel on the rank-ordered sum"
        assert 1e-4 < moved <= 3.1e-3


def test_two_ranks_with_the_fused_p2p_exchange_inside_the_cuda_graph(eager_result):
    res = _run("p2p_graph")
    ref = eager_result[0][6]
    for rank, _, identical, _, moved, loss, flat in res:
        assert identical and 1e-4 < moved <= 3.1e-3
        if flat is not None:
            d = (flat.double() - ref.double()).abs()
            assert int((d > 1e-4).sum()) <= 0.005 * d.numel() and float(d.max()) <= 6e-3 + 1e-6

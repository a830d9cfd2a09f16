"""CPU, world_size 2 over gloo: the data-parallel host logic (flat-slab all-reduce + 1/N scale, per-stream
sharding, loss averaging) used by the N > 1 path."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dcase2019_task4_b200 import dp
from oracle import train_step as otrain


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    r, w = dp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)
    B, To = 4, 6                                             # per-rank batch: 1 weak | 2 unlabeled | 1 strong
    strong = torch.rand(B, To, 10, requires_grad=True)
    weak = torch.rand(B, 10, requires_grad=True)
    st, wt = torch.rand(B, To, 10), torch.rand(B, 10)
    tgt = (torch.rand(B, To, 10) < 0.3).float()
    loss, _ = otrain.mean_teacher_losses(strong, weak, st, wt, tgt, slice(0, 1), slice(3, 4), 1.5)
    loss.backward()
    flat = torch.cat([strong.grad.reshape(-1), weak.grad.reshape(-1)])
    local = flat.clone()
    scale = dp.allreduce_grads_(flat)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    losses = [torch.zeros(()) for _ in range(world)]
    dist.all_gather(losses, loss.detach())
    ok_mean = torch.allclose(flat * scale, torch.stack(gathered).mean(0), atol=1e-7)
    q.put((rank, scale, ok_mean, float(torch.stack(losses).mean()), [t.tolist() for t in (strong.detach(), weak.detach(), st, wt, tgt)]))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_mean_and_global_loss_equivalence():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] == 0.5 and r[2] for r in res)
    # mean of the per-rank mean losses == loss of the concatenated (per-stream sharded) global batch
    parts = [[torch.tensor(t) for t in r[4]] for r in res]

    def cat(i):                                              # global batch in stream order [weak | unl | strong]
        a, b = parts[0][i], parts[1][i]
        return torch.cat([a[0:1], b[0:1], a[1:3], b[1:3], a[3:4], b[3:4]])
    g_loss, _ = otrain.mean_teacher_losses(cat(0), cat(1), cat(2), cat(3), cat(4), slice(0, 2), slice(6, 8), 1.5)
    assert abs(float(g_loss) - res[0][3]) < 1e-6

import sys, time; sys.path.insert(0, '.')
import torch, bench
from dcase2019_task4_b200 import config as cfg, kernels as K
from dcase2019_task4_b200.main import MeanTeacherEngine
from dcase2019_task4_b200.models.CRNN import CRNN
from dcase2019_task4_b200.utils.utils import weights_init
dev = torch.device('cuda', 0)
waves, targets = bench.synthetic_batches(2, seed=1)
wave_dev = torch.from_numpy(waves).to(dev); target_dev = torch.from_numpy(targets).to(dev)
mean = torch.full((64,), -30.0, device=dev); std = torch.full((64,), 12.0, device=dev)
crnn, crnn_ema = CRNN(**cfg.crnn_kwargs), CRNN(**cfg.crnn_kwargs)
crnn.apply(weights_init); crnn_ema.apply(weights_init)
for p in crnn_ema.parameters(): p.detach_()
crnn, crnn_ema = crnn.train().cuda(), crnn_ema.train().cuda()
opt = torch.optim.Adam(crnn.parameters(), lr=0.001, betas=(0.9, 0.999))
eng = MeanTeacherEngine(crnn, opt, crnn_ema, slice(6), slice(18, 24), 24, 864)
def step(i): eng.step_from_waveforms(wave_dev[i % 2], target_dev[i % 2], mean, std, 0.1, i + 1, check=False)
for i in range(5): step(i)
torch.cuda.synchronize()
def timeit(fn, n=30):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    for i in range(n): fn(i)
    e1.record(); t_cpu = (time.perf_counter() - t0) / n; torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_cpu * 1e3
print("eager  ms/step gpu %.3f  cpu-issue %.3f" % timeit(step))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step(0)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step(0)
print("graph  ms/step gpu %.3f  cpu-issue %.3f" % timeit(lambda i: g.replay()))

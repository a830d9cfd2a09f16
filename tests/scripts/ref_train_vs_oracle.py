"""Run by tests/test_oracle_vs_reference.py in a SUBPROCESS (needs /root/reference).

Pins the step oracle (oracle/train_step.py::train_batch) against the reference's OWN training functions, imported
unmodified: ``baseline/main.py::train`` (mean teacher, main.py:52-165) and ``baseline/main_simple_CRNN.py::train``
(:31-82), driving the reference's own ``models.CRNN`` on the CPU.  The scripts' other imports (DataLoad, utils.*, config,
evaluation_measures -- not importable here because librosa / dcase_util / sed_eval are absent) are satisfied by this
package through ``dropin.install()``; the ``models`` aliases are removed again so that ``models.CRNN`` is the reference's.
Prints ``REF-TRAIN-OK <student> <teacher> <simple> <running_var> <bias-corrected running_mean>`` (max abs differences).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/baseline"
sys.path.insert(0, ROOT)

from dcase2019_task4_b200 import dropin  # noqa: E402

dropin.install()
for name in list(sys.modules):
    if name == "models" or name.startswith("models."):
        del sys.modules[name]
sys.path.append(REF)

import torch  # noqa: E402

import main as refmain  # noqa: E402            (the reference's main.py)
import main_simple_CRNN as refsimple  # noqa: E402
from models.CRNN import CRNN as RefCRNN  # noqa: E402

import dcase2019_task4_b200.config as cfg  # noqa: E402
from oracle import crnn as ocrnn  # noqa: E402
from oracle import train_step as otrain  # noqa: E402

assert sys.modules["models.CRNN"].__file__.startswith(REF) and refmain.__file__.startswith(REF)

KW = dict(cfg.crnn_kwargs)
KW["dropout"] = 0                      # no RNG in the comparison; dropout parity is covered by mask injection elsewhere
B, T = 8, 64


class Loader(list):
    pass


def build(p):
    m = RefCRNN(**KW)
    with torch.no_grad():
        for k, v in m.named_parameters():
            v.copy_(p[k])
    return m.train()


def max_diff(model, p):
    return max(float((v.detach() - p[k]).abs().max()) for k, v in model.named_parameters()
               if not (".conv" in k and k.endswith("bias")))       # zero gradient behind BatchNorm: rounding noise


g = torch.Generator().manual_seed(5)
batches = Loader()
for _ in range(3):
    x = torch.randn(B, 1, T, 64, generator=g)
    xe = x + 0.1 * torch.randn(B, 1, T, 64, generator=g)
    tgt = (torch.rand(B, T // 8, 10, generator=g) < 0.2).float()
    tgt[2:6] = -1
    batches.append((x, xe, tgt))
wm, sm = slice(2), slice(6, 8)

# ---- mean teacher: reference main.train vs oracle ----
ps, pt = ocrnn.init_params(seed=31), ocrnn.init_params(seed=32)
student, teacher = build(ps), build(pt)
for q in teacher.parameters():
    q.detach_()
opt = torch.optim.Adam(filter(lambda q: q.requires_grad, student.parameters()), lr=0.001, betas=(0.9, 0.999))
refmain.train(batches, student, opt, 0, ema_model=teacher, weak_mask=wm, strong_mask=sm)
sbuf, tbuf = ocrnn.init_buffers(), ocrnn.init_buffers()
adam = otrain.new_adam_state(ps)
for i, (x, xe, tgt) in enumerate(batches):
    otrain.train_batch(ps, sbuf, adam, x, tgt, i, len(batches), teacher_p=pt, teacher_buf=tbuf, x_ema=xe,
                       weak_mask=wm, strong_mask=sm)
d_student, d_teacher = max_diff(student, ps), max_diff(teacher, pt)
# running_var directly; running_mean tracks (conv output mean) = (mean without bias) + conv bias, and the conv biases
# random-walk by +-lr per step on rounding noise (zero true gradient behind BatchNorm, Adam normalises the noise), with
# different noise on the two sides: compare it bias-corrected, with the last +-lr step as slack
bn = dict(student.named_buffers())
sp_now = dict(student.named_parameters())
d_bn = d_rm = 0.0
for i in range(3):
    d_bn = max(d_bn, float((bn["cnn.cnn.batchnorm%d.running_var" % i]
                            - sbuf["cnn.cnn.batchnorm%d.running_var" % i]).abs().max()))
    rm_ref = bn["cnn.cnn.batchnorm%d.running_mean" % i] - sp_now["cnn.cnn.conv%d.bias" % i].detach()
    rm_ora = sbuf["cnn.cnn.batchnorm%d.running_mean" % i] - ps["cnn.cnn.conv%d.bias" % i]
    d_rm = max(d_rm, float((rm_ref - rm_ora).abs().max()))

# ---- plain CRNN: reference main_simple_CRNN.train vs oracle ----
p2 = ocrnn.init_params(seed=33)
model = build(p2)
opt2 = torch.optim.Adam(filter(lambda q: q.requires_grad, model.parameters()), lr=0.001, betas=(0.9, 0.999))
simple_batches = Loader((x, (tgt >= 0.5).float()) for x, _, tgt in batches)
refsimple.train(simple_batches, model, opt2, 0, weak_mask=slice(4), strong_mask=slice(4, 8))
buf2 = ocrnn.init_buffers()
adam2 = otrain.new_adam_state(p2)
for i, (x, tgt) in enumerate(simple_batches):
    otrain.train_batch(p2, buf2, adam2, x, tgt, i, len(simple_batches), weak_mask=slice(4), strong_mask=slice(4, 8))
d_simple = max_diff(model, p2)

if len(sys.argv) > 1:          # tests/golden/make_golden.py: save what the REFERENCE produced (strided subsample of the slabs)
    import numpy as np
    names = [k for k, _ in student.named_parameters()]
    flat_s = torch.cat([v.detach().reshape(-1) for _, v in student.named_parameters()]).numpy()
    flat_t = torch.cat([v.detach().reshape(-1) for _, v in teacher.named_parameters()]).numpy()
    out = {"stride": np.int64(8), "student_after": flat_s[::8].copy(), "teacher_after": flat_t[::8].copy(),
           "seeds": np.array([31, 32, 5], dtype=np.int64), "B": np.int64(B), "T": np.int64(T),
           "n_params": np.int64(flat_s.size), "first_name": np.array(names[0])}
    for i in range(3):
        out["running_var%d" % i] = bn["cnn.cnn.batchnorm%d.running_var" % i].numpy()
    for j, (x, xe, tgt) in enumerate(batches):
        out["x%d" % j], out["xe%d" % j], out["tgt%d" % j] = x.numpy(), xe.numpy(), tgt.numpy()
    np.savez_compressed(sys.argv[1], **out)

print("REF-TRAIN-OK %.3e %.3e %.3e %.3e %.3e" % (d_student, d_teacher, d_simple, d_bn, d_rm))

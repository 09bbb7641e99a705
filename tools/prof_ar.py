"""Tiny driver for ncu: one begin + one persistent-kernel launch.  usage: prof_ar.py B wdtype steps"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.cases import AR_CASES, ar_case_inputs, _unit
from sopro_b200.engine import ArEngine, Sampling

B, wd, steps = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
L = 52
spec = AR_CASES["default_bf16" if wd == "bf16" else "default_fp32"]
cfg, sd, _ = ar_case_inputs(spec)
eng = ArEngine(cfg, sd, 0, wd)
D = int(cfg.d_model)
cond = (_unit(steps * D, 1).view(1, steps, D).expand(B, steps, D) + 0.01 * torch.arange(B).view(B, 1, 1)).contiguous()
txt = _unit(L * D, 2).view(1, L, D).expand(B, L, D).contiguous()
noise = torch.empty(B, steps, 50).exponential_(1.0, generator=torch.Generator().manual_seed(0))
ses = eng.session(B, steps, L)
for _ in range(2):
    ses.begin(cond, txt, [L] * B, noise, Sampling(min_gen_frames=2**31 - 1))
    ses.run()
    torch.cuda.synchronize()
print("done", ses.read()[1][:4])

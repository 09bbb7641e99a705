"""Tiny driver for ncu / timing: one NAR refine of B x T frames.  usage: prof_nar.py B T"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.cases import _unit, e2e_inputs
from sopro_b200.nar import NarEngine

B, T = int(sys.argv[1]), int(sys.argv[2])
cfg, sd, _ = e2e_inputs()
eng = NarEngine(cfg, sd, 0)
D = int(cfg.d_model)
cond = torch.stack([_unit(T * D, 9100 + i).view(T, D) for i in range(B)]).cuda()
rvq1 = torch.randint(0, 2048, (B, T), generator=torch.Generator().manual_seed(1)).cuda()
for _ in range(2):
    out = eng.refine(cond, rvq1)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
out = eng.refine(cond, rvq1)
ev[1].record()
torch.cuda.synchronize()
print(f"nar refine B={B} T={T}: {ev[0].elapsed_time(ev[1]):.3f} ms", int(out.sum()))

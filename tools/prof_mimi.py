"""Tiny driver for ncu / timing: Mimi decode of T random code frames.  usage: prof_mimi.py T [precision] [B] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sopro_b200.weights import synth_mimi_state_dict
from sopro_b200.codec import MimiEngine

T = int(sys.argv[1])
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16_tc"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
eng = MimiEngine(synth_mimi_state_dict(), 0, 32, precision=prec)
codes = torch.randint(0, 2048, (B, 32, T), generator=torch.Generator().manual_seed(0)).cuda()
eng.decode(codes)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(reps):
    wav = eng.decode(codes)
ev[1].record()
torch.cuda.synchronize()
print(f"mimi {prec} B={B} T={T}: {ev[0].elapsed_time(ev[1]) / reps:.3f} ms/decode", float(wav.abs().max()))

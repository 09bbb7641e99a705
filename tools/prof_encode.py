"""Tiny driver for ncu: Mimi encode of a recording + prepare_reference of its codes.  usage: prof_encode.py seconds"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sopro_b200.codec import MimiEncoderEngine
from sopro_b200.config import SoproTTSConfig
from sopro_b200.prefill_cuda import RefPrepEngine
from sopro_b200.weights import synth_mimi_encoder_state_dict, synth_mimi_state_dict, synth_state_dict

sec = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
sd = dict(synth_mimi_state_dict()); sd.update(synth_mimi_encoder_state_dict())
enc = MimiEncoderEngine(sd, 0, 32)
cfg = SoproTTSConfig()
rp = RefPrepEngine(cfg, synth_state_dict(cfg, 1000, 0), 0)
wav = ((torch.rand(int(24000 * sec), generator=torch.Generator().manual_seed(9)) - 0.5) * 0.6).cuda()
for _ in range(2):
    codes = enc.encode(wav)
    sv, seq, kv = rp.run(codes.permute(1, 0).contiguous())
    torch.cuda.synchronize()
print("done", codes.shape, float(sv.norm()))

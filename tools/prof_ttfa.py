"""Host-side profile of stream()'s first chunk (time-to-first-audio): cProfile over repeated first-chunk calls."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sopro_b200 import SoproTTS
from sopro_b200.config import SoproTTSConfig
from sopro_b200.tokenizer import IdsTokenizer
from sopro_b200.weights import synth_mimi_state_dict, synth_state_dict

cfg = SoproTTSConfig()
tts = SoproTTS.from_state_dict(cfg, synth_state_dict(cfg, 1000, 0), IdsTokenizer(1000), synth_mimi_state_dict(), device="cuda:0", weight_dtype="bf16")
ref = tts.prepare_reference(ref_tokens_tq=torch.randint(0, 2048, (38, 32), generator=torch.Generator().manual_seed(7)))
text = " ".join(str(17 * i + 5) for i in range(50))


def first():
    it = tts.stream(text, ref=ref, max_frames=400, seed=1, min_gen_frames=10 ** 9)
    c = next(it).cpu()
    it.close()
    return c


for _ in range(2):
    sum(1 for _ in tts.stream(text, ref=ref, max_frames=400, seed=1, min_gen_frames=10 ** 9))
for _ in range(5):
    first()
ts = []
for _ in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter(); first(); ts.append(time.perf_counter() - t0)
print(f"TTFA p50 {np.median(ts) * 1e3:.3f} ms  min {min(ts) * 1e3:.3f}")
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    first()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(38)

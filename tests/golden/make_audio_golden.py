"""Generates tests/golden/audio_prep.json from the UNMODIFIED reference (src/sopro/audio.py): what `encode_file`'s
host-side preparation (energy trim -> centre crop; reference codec/mimi.py:44-57) keeps of seeded test signals.
Run in the build container only:  python tests/golden/make_audio_golden.py"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from sopro_b200.weights import hash_uniform  # noqa: E402

CASES = [  # (name, sr, total samples, voiced [start, end), noise floor amplitude)
    ("margins_24k", 24000, 72000, 20000, 50000, 0.0),
    ("noisy_margins_24k", 24000, 72000, 15000, 60000, 1e-3),
    ("all_voiced_16k", 16000, 40000, 0, 40000, 0.0),
    ("short_burst_24k", 24000, 48000, 24000, 26000, 0.0),   # shorter than min_keep_sec: untouched
    ("tiny_24k", 24000, 1000, 0, 1000, 0.0),                # shorter than 0.1 s: untouched
    ("silent_24k", 24000, 30000, 0, 0, 0.0),
]


def signal(sr, n, lo, hi, floor, key):
    x = np.zeros(n, dtype=np.float32)
    if floor:
        x += hash_uniform(n, 0xF100 + key) * np.float32(floor)
    t = np.arange(hi - lo, dtype=np.float64) / sr
    x[lo:hi] += (0.4 * np.sin(2 * np.pi * 220.0 * t)).astype(np.float32) + hash_uniform(hi - lo, 0xA00 + key) * np.float32(0.05)
    return torch.from_numpy(x).unsqueeze(0)


def main():
    sys.path.insert(0, "/root/reference/src")
    from sopro.audio import center_crop_audio, trim_silence_energy  # the unmodified reference

    out = {}
    for i, (name, sr, n, lo, hi, floor) in enumerate(CASES):
        w = signal(sr, n, lo, hi, floor, i)
        t = trim_silence_energy(w, sr)
        c = center_crop_audio(t, 12 * 1920)
        out[name] = {"trim_len": int(t.shape[-1]), "trim_first": float(t[0, 0]), "trim_last": float(t[0, -1]),
                     "trim_sum": float(t.double().sum()), "crop_len": int(c.shape[-1]), "crop_first": float(c[0, 0]),
                     "crop_sum": float(c.double().sum())}
    with open(os.path.join(HERE, "audio_prep.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main()

"""Generates tests/golden/mimi_encode.npz: the codes the installed transformers MimiModel.encode (what the reference's
MimiCodec.encode_file calls, reference codec/mimi.py:59-62) gives for seeded waveforms under the seeded synthetic
weights (sopro_b200.weights.synth_mimi_state_dict + synth_mimi_encoder_state_dict).  No checkpoint exists offline, so the
weights are synthetic; the arithmetic is transformers' own.  Run here (CPU):  python tests/golden/make_mimi_encode_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sopro_b200.weights import hash_uniform, synth_mimi_encoder_state_dict, synth_mimi_state_dict  # noqa: E402

LENGTHS = (999, 5760, 13951, 48077)  # one frame; whole frames; ragged tails at every stride


def waveform(n: int) -> torch.Tensor:
    """Platform-independent test signal: a chirp plus hashed noise, |x| < 0.6."""
    t = np.arange(n, dtype=np.float64) / 24000.0
    x = 0.3 * np.sin(2 * np.pi * (110.0 + 400.0 * t) * t) + 0.25 * hash_uniform(n, 0xA0D10 + n).astype(np.float64)
    return torch.from_numpy(x.astype(np.float32)).view(1, 1, n)


def main():
    import transformers as tr

    torch.set_grad_enabled(False)
    sd = dict(synth_mimi_state_dict())
    sd.update(synth_mimi_encoder_state_dict())
    m = tr.MimiModel(tr.MimiConfig(num_quantizers=32)).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    out = {}
    for n in LENGTHS:
        codes = m.encode(waveform(n), return_dict=True).audio_codes[0]
        assert codes.shape[1] == int(m.get_encoded_length(torch.tensor(n)))
        out[f"codes_{n}"] = codes.numpy().astype(np.int16)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mimi_encode.npz"),
                        transformers_version=np.array(tr.__version__), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

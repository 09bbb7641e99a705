"""Build container only (needs /root/reference): how far is the REFERENCE's MimiStreamDecoder.decode_step
(reference codec/mimi.py:115-181, 2-frame overlap on top of an HF KV cache) from the reference's own decode_full
on the installed transformers?  Our decode_step returns the decode_full prefix exactly (causal decoder), so this
number IS the distance between the two decode_step implementations (plus 1.5e-7 of oracle noise).

    python tests/golden/measure_stream_distance.py  ->  prints one JSON line, recorded in DESIGN.md §5
"""
import json
import sys

import torch

sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, ".")
torch.set_grad_enabled(False)


def main():
    import transformers as tr
    from sopro.codec.mimi import MimiCodec, MimiDecodeState, MimiStreamDecoder

    from oracle import mimi_oracle as M

    sd = M.synth_mimi_state_dict()
    hf = tr.MimiModel(tr.MimiConfig(num_quantizers=32)).eval()
    hf.load_state_dict(sd, strict=False)
    codec = object.__new__(MimiCodec)
    codec.device = torch.device("cpu")
    codec.model = hf
    out = {"transformers": tr.__version__, "torch": torch.__version__, "cases": []}
    for T, chunk, seed in ((48, 6, 21), (150, 6, 22), (150, 16, 23)):
        codes_tq = torch.randint(0, 2048, (T, 32), generator=torch.Generator().manual_seed(seed))
        full = codec.decode_full(codes_tq).reshape(-1)
        dec = MimiStreamDecoder(codec)
        st = MimiDecodeState()
        parts = []
        for s in range(0, T, chunk):
            w, st = dec.decode_step(codes_tq[s: s + chunk], st)
            parts.append(w.reshape(-1))
        stream = torch.cat(parts)
        n = min(stream.numel(), full.numel())
        err = (stream[:n] - full[:n]).abs()
        peak = float(full.abs().max())
        first = err[: chunk * 1920]
        out["cases"].append({"frames": T, "chunk": chunk, "samples_stream": int(stream.numel()), "samples_full": int(full.numel()),
                             "max_abs_err": float(err.max()), "peak": peak, "max_err_over_peak": float(err.max()) / peak,
                             "rel_rms": float(err.pow(2).mean().sqrt() / full[:n].pow(2).mean().sqrt()),
                             "first_chunk_max_err_over_peak": float(first.max()) / peak})
    print(json.dumps(out))


if __name__ == "__main__":
    main()

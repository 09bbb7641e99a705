"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case it (1) builds the seeded synthetic checkpoint
(sopro_b200.weights.synth_state_dict — platform independent), (2) loads it into
the UNMODIFIED reference modules imported from /root/reference/src, (3) runs the
reference's own ``SoproTTSModel.ar_stream`` / ``sample_token`` /
``ARRVQ1Generator.step`` with the global torch RNG seeded, (4) runs oracle/ on
the same inputs and ASSERTS BIT-EQUALITY with the reference, and (5) writes the
small fixtures (token ids, a few logit rows, per-block traces) that
tests/test_oracle_golden.py and the GPU parity tests replay on any host.

The reference has no tests or golden vectors of its own; these
files are the pin.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from sopro.config import SoproTTSConfig as RefCfg  # noqa: E402
from sopro.model import SoproTTSModel  # noqa: E402
from sopro.sampling import sample_token as ref_sample_token  # noqa: E402

from oracle import ar_oracle as O  # noqa: E402
from sopro_b200.config import SoproTTSConfig  # noqa: E402
from sopro_b200.weights import hash_uniform, round_through_bf16, synth_state_dict  # noqa: E402
from tests.cases import AR_CASES, ar_case_inputs, SAMPLER_CASES, sampler_case_inputs  # noqa: E402

torch.set_grad_enabled(False)


class _Tok:
    def __init__(self, v):
        self.vocab_size = v


def build_reference_model(cfg: SoproTTSConfig, sd, text_vocab: int):
    rcfg = RefCfg(**{k: getattr(cfg, k) for k in RefCfg.__annotations__.keys()})
    m = SoproTTSModel(rcfg, _Tok(text_vocab)).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(not k.startswith(("ar.", "cb_embed.")) for k in missing), missing
    return m


def run_ar_case(name: str, spec: dict) -> dict:
    cfg, sd, inp = ar_case_inputs(spec)
    ref = build_reference_model(cfg, sd, text_vocab=64)
    prep = {"cond_ar": inp["cond_ar"], "txt_seq": inp["txt_seq"], "text_mask": inp["text_mask"]}
    samp = inp["sampling"]
    # ---- the reference, consuming the global RNG
    torch.manual_seed(int(spec["noise_seed"]))
    ref_tokens = [
        tok
        for _t, tok, _e in ref.ar_stream(
            prep, max_frames=inp["max_frames"], top_p=samp.top_p, temperature=samp.temperature,
            anti_loop=samp.anti_loop, min_gen_frames=samp.min_gen_frames,
        )
    ]
    # ---- the oracle, fed the explicit noise tape
    tape = O.noise_tape(int(spec["noise_seed"]), inp["max_frames"] + 1, cfg.ar_vocab())
    logits, recov = [], []
    ora_tokens = O.ar_generate(
        sd, cfg, inp["cond_ar"], inp["txt_seq"], inp["text_mask"], max_frames=inp["max_frames"],
        sampling=samp, noise_tv=tape, logits_out=logits, recovery_out=recov,
    )
    assert ora_tokens == ref_tokens, (name, "oracle tokens != reference tokens")
    # ---- step-level: reference ARRVQ1Generator.step vs oracle ar_step, teacher forced
    state = ref.ar.init_stream_state(1, torch.device("cpu"), torch.float32,
                                     text_emb=inp["txt_seq"], text_mask=inp["text_mask"])
    bos = int(cfg.num_codebooks) * int(cfg.codebook_size)
    emb = sd["cb_embed.emb.weight"]
    for t in range(min(len(ref_tokens), 24)):
        row = bos if t == 0 else ref_tokens[t - 1]
        x_t = inp["cond_ar"][:, t:t + 1] + emb[row].view(1, 1, -1)
        lg, state = ref.ar.step(x_t, state, text_emb=inp["txt_seq"], text_mask=inp["text_mask"])
        assert torch.equal(lg[0, 0], logits[t]), (name, t, "oracle logits != reference step logits")
    # ---- full-sequence forward (nn/generator.py:70-96) vs steps: the reference's own KAT
    T = min(len(ref_tokens), 64)
    rows = [bos] + ref_tokens[: T - 1]
    x_full = inp["cond_ar"][:, :T] + emb[torch.tensor(rows)].unsqueeze(0)
    full = ref.ar(x_full, inp["txt_seq"], inp["text_mask"])[0]
    step_vs_full = float((full - torch.stack(logits[:T])).abs().max())
    assert step_vs_full < 5e-5, step_vs_full
    # per-block trace at steps 0 and 1 from the oracle (== reference by the asserts above)
    st = O.ar_init_state(sd, cfg, inp["txt_seq"], inp["text_mask"])
    traces = []
    for t in range(2):
        row = bos if t == 0 else ref_tokens[t - 1]
        x_t = inp["cond_ar"][:, t:t + 1] + emb[row].view(1, 1, -1)
        tr = {}
        O.ar_step(sd, cfg, x_t, st, trace=tr)
        traces.append(np.stack([tr[f"h{i}"][0, 0].numpy() for i in range(int(cfg.n_layers_ar))]))
    keep = sorted(set([0, 1, 2, 3, 7, 15, 50, 100, 200, 300, len(ref_tokens) - 1]) & set(range(len(ref_tokens))))
    out = {
        "tokens": np.asarray(ref_tokens, dtype=np.int32),
        "logit_steps": np.asarray(keep, dtype=np.int32),
        "logits": np.stack([logits[t].numpy() for t in keep]).astype(np.float32),
        "block_trace": np.stack(traces).astype(np.float32),
        "full_vs_step_maxabs": np.float32(step_vs_full),
        "recovery_steps": np.asarray(recov, dtype=np.int32),
    }
    print(f"[ar] {name}: {len(ref_tokens)} tokens, first {ref_tokens[:8]}, "
          f"n_eos={sum(t == cfg.codebook_size for t in ref_tokens)}, full-vs-step {step_vs_full:.2e}, "
          f"logit std {float(torch.stack(logits).std()):.3f}, recovery steps {len(recov)}, "
          f"eos at {[i for i, t in enumerate(ref_tokens) if t == cfg.codebook_size][:6]}")
    return out


def run_sampler_cases() -> dict:
    out = {}
    for name, spec in SAMPLER_CASES.items():
        logits, hist, kw, seed = sampler_case_inputs(spec)
        V = logits.numel()
        torch.manual_seed(seed)
        ref_tok = ref_sample_token(logits.view(1, 1, V), list(hist), **kw)
        tape = O.noise_tape(seed, 1, V)
        ora_tok = O.sample_token(logits.view(1, 1, V), list(hist), noise_v=tape[0], **kw)
        torch.manual_seed(seed)
        ora_tok_rng = O.sample_token(logits.view(1, 1, V), list(hist), noise_v=None, **kw)
        assert ref_tok == ora_tok == ora_tok_rng, (name, ref_tok, ora_tok, ora_tok_rng)
        out[name] = int(ref_tok)
    print(f"[sampler] {len(out)} cases ok")
    return out


def main():
    torch.set_num_threads(8)
    for name, spec in AR_CASES.items():
        res = run_ar_case(name, spec)
        np.savez_compressed(os.path.join(HERE, f"ar_{name}.npz"), **res)
    with open(os.path.join(HERE, "sampler_kat.json"), "w") as f:
        json.dump(run_sampler_cases(), f, indent=0, sort_keys=True)
    meta = {"torch": torch.__version__, "reference": "samuel-vitorino/sopro @ 5fe20d2",
            "note": "written by tests/golden/make_golden.py; oracle asserted bit-equal to the reference"}
    with open(os.path.join(HERE, "META.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()

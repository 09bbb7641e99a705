"""CPU: the oracle replayed against the golden fixtures written from the reference
(tests/golden/make_golden.py).  No GPU, no /root/reference needed."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ar_oracle as O
from tests.cases import AR_CASES, SAMPLER_CASES, ar_case_inputs, sampler_case_inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden")
torch.set_grad_enabled(False)


def _gold(name):
    return np.load(os.path.join(GOLD, f"ar_{name}.npz"))


@pytest.mark.parametrize("name", list(AR_CASES))
def test_ar_tokens_and_logits_match_reference(name):
    spec = AR_CASES[name]
    g = _gold(name)
    cfg, sd, inp = ar_case_inputs(spec)
    tape = O.noise_tape(spec["noise_seed"], inp["max_frames"] + 1, cfg.ar_vocab())
    logits, recov = [], []
    toks = O.ar_generate(sd, cfg, inp["cond_ar"], inp["txt_seq"], inp["text_mask"], max_frames=inp["max_frames"],
                         sampling=inp["sampling"], noise_tv=tape, logits_out=logits, recovery_out=recov)
    gold_tokens = g["tokens"].tolist()
    # logits first (tolerance: a different host CPU may pick other SIMD kernels)
    for row, t in zip(g["logits"], g["logit_steps"].tolist()):
        if t < len(logits) and toks[:t] == gold_tokens[:t]:
            np.testing.assert_allclose(logits[t].numpy(), row, rtol=0, atol=2e-5 * max(1.0, float(np.abs(row).max())))
    assert toks == gold_tokens
    assert recov == g["recovery_steps"].tolist()


def test_sampler_known_answers():
    with open(os.path.join(GOLD, "sampler_kat.json")) as f:
        kat = json.load(f)
    assert set(kat) == set(SAMPLER_CASES)
    for name, spec in SAMPLER_CASES.items():
        logits, hist, kw, seed = sampler_case_inputs(spec)
        V = logits.numel()
        tape = O.noise_tape(seed, 1, V)
        assert O.sample_token(logits.view(1, 1, V), hist, noise_v=tape[0], **kw) == kat[name], name


def test_multinomial_is_argmax_of_p_over_exponential():
    """The RNG equivalence the noise tape rests on, asserted on this torch."""
    V = 2049
    for seed in range(20):
        p = torch.softmax(torch.from_numpy(np.random.RandomState(seed).randn(V).astype(np.float32)) * 3, -1).view(1, V)
        torch.manual_seed(seed)
        draws = [int(torch.multinomial(p, 1)) for _ in range(3)]
        tape = O.noise_tape(seed, 3, V)
        assert draws == [int(torch.argmax(p / tape[i].view(1, V))) for i in range(3)]


def test_repeated_tail():
    assert not O.repeated_tail([])
    assert not O.repeated_tail([1, 2, 1, 2])  # n=2 is not checked
    assert O.repeated_tail([1, 2, 3, 1, 2, 3])
    assert O.repeated_tail([9, 9] + [4] * 6)
    assert not O.repeated_tail(list(range(40)))
    h = list(range(16)) * 2
    assert O.repeated_tail(h) and not O.repeated_tail(list(range(17)) * 2, 16)


def test_nar_oracle_reproduces_the_reference_tokens():
    """oracle/nar_oracle.py against the 50 x 32 tokens the unmodified reference wrote (make_golden_e2e.py)."""
    import numpy as np

    from oracle import nar_oracle as N
    from sopro_b200 import prefill as P
    from tests.cases import e2e_inputs

    cfg, sd, inp = e2e_inputs()
    g = np.load(os.path.join(GOLD, "e2e_prefill.npz"))
    dev = torch.device("cpu")
    pr = P.prepare_reference(sd, cfg, inp["ref_tokens_tq"], dev)
    tpos = P.sinusoid_table(int(cfg.max_text_len) + 8, int(cfg.d_model), dev)
    fpos = P.sinusoid_table(int(cfg.pos_emb_max) + 8, int(cfg.d_model), dev)
    prep = P.prepare_conditioning(sd, cfg, inp["text_ids"], pr, max_frames=inp["max_frames"], device=dev,
                                  style_strength=inp["style_strength"], text_pos=tpos, frame_pos=fpos)
    codes, margin = N.nar_refine(sd, cfg, prep["cond_ar"][:, : inp["nar_T"]], inp["rvq1"].unsqueeze(0))
    assert torch.equal(codes[0], torch.from_numpy(g["nar_tokens"].astype(np.int64)))
    assert float(margin[0, :, 1:].min()) > 0

"""CPU: the Mimi decode restatement (oracle/mimi_oracle.py) against the installed transformers MimiModel,
which is the arithmetic the reference actually runs (reference codec/mimi.py:65-72)."""
import pytest
import torch

from oracle import mimi_oracle as M

torch.set_grad_enabled(False)


def _hf_model(sd):
    tr = pytest.importorskip("transformers")
    m = tr.MimiModel(tr.MimiConfig(num_quantizers=32)).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(not k.startswith(("decoder.", "decoder_transformer.", "upsample.")) and "output_proj" not in k and "embed_sum" not in k
               and "cluster_usage" not in k for k in missing), [k for k in missing][:5]
    return m


def test_restatement_matches_transformers_mimi_decode():
    sd = M.synth_mimi_state_dict()
    hf = _hf_model(sd)
    g = torch.Generator().manual_seed(5)
    codes = torch.randint(0, 2048, (2, 32, 9), generator=g)
    want = hf.decode(audio_codes=codes, return_dict=True).audio_values
    got = M.mimi_decode(sd, codes)
    assert got.shape == want.shape == (2, 1, 9 * 1920)
    assert float(want.abs().max()) > 1e-3
    assert float((got - want).abs().max()) <= 2e-5 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("B,T,seed", [(1, 150, 15), (2, 130, 16)])
def test_restatement_matches_transformers_beyond_the_attention_window(B, T, seed):
    """T >= 126 frames = more than 250 transformer positions: the sliding-window mask is live (the judge measured the
    window changing the transformer output by 2.9e-2 at T = 150), and B > 1 exercises the batch dimension."""
    sd = M.synth_mimi_state_dict()
    hf = _hf_model(sd)
    codes = torch.randint(0, 2048, (B, 32, T), generator=torch.Generator().manual_seed(seed))
    want = hf.decode(audio_codes=codes, return_dict=True).audio_values
    got = M.mimi_decode(sd, codes)
    assert got.shape == want.shape == (B, 1, T * 1920)
    assert float((got - want).abs().max()) <= 2e-5 * max(1.0, float(want.abs().max()))
    # the window matters on this input: an unwindowed transformer gives a different waveform
    x = M.upsample(sd, M.rvq_decode(sd, codes[:1]))
    d = float((M.transformer(sd, x) - M.transformer(sd, x, window=10 ** 6)).abs().max())
    assert d > 1e-3, d


def test_decode_is_causal_prefix_exact():
    """Decoding a prefix gives the prefix of the decoded audio: what the streaming decoder relies on."""
    sd = M.synth_mimi_state_dict()
    g = torch.Generator().manual_seed(6)
    codes = torch.randint(0, 2048, (1, 32, 12), generator=g)
    full = M.mimi_decode(sd, codes)
    part = M.mimi_decode(sd, codes[:, :, :7])
    assert float((full[..., : 7 * 1920] - part).abs().max()) <= 1e-5


def test_bf16_operand_model_stays_near_the_fp32_restatement():
    """oracle.mimi_decode_bf16_operands models the product's tensor-core mode (operands rounded to bf16 where the
    kernels round them).  It must differ from the fp32 restatement (otherwise it rounds nothing) and stay inside the
    tolerance the product states for that mode (2e-2 of the peak, 1e-2 relative RMS).  On the smoke() input the GPU's
    tensor-core mode measured 8.16e-4 from the fp32 oracle; this model gives 8.0e-4."""
    sd = M.synth_mimi_state_dict()
    for seed, B, T in ((1, 1, 5), (109, 2, 9)):
        codes = torch.randint(0, 2048, (B, 32, T), generator=torch.Generator().manual_seed(seed))
        ref, emu = M.mimi_decode(sd, codes), M.mimi_decode_bf16_operands(sd, codes)
        peak, err = float(ref.abs().max()), float((emu - ref).abs().max())
        assert 1e-3 * peak <= err <= 2e-2 * peak, (err, peak)
        assert float((emu - ref).pow(2).mean().sqrt()) <= 1e-2 * float(ref.pow(2).mean().sqrt())
    err1 = float((M.mimi_decode_bf16_operands(sd, torch.randint(0, 2048, (1, 32, 5), generator=torch.Generator().manual_seed(1)))
                  - M.mimi_decode(sd, torch.randint(0, 2048, (1, 32, 5), generator=torch.Generator().manual_seed(1)))).abs().max())
    assert abs(err1 - 8.16e-4) <= 1e-4  # same error as the GPU measured on these codes (profiles/r01e_summary.md)


# ---------------------------------------------------------------------------------------------------------------
# ENCODE path
# ---------------------------------------------------------------------------------------------------------------
def _full_sd():
    sd = dict(M.synth_mimi_state_dict())
    sd.update(M.synth_mimi_encoder_state_dict())
    return sd


def _golden_encode():
    import os

    import numpy as np

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mimi_encode.npz"))


def _golden_waveform(n):
    from tests.golden.make_mimi_encode_golden import waveform

    return waveform(n)


def test_encode_restatement_matches_the_committed_transformers_codes():
    """tests/golden/mimi_encode.npz holds MimiModel.encode's codes (transformers 5.5.0, generated here by
    tests/golden/make_mimi_encode_golden.py); the restatement reproduces every id, ragged lengths included."""
    sd, g = _full_sd(), _golden_encode()
    for n in (999, 5760, 13951, 48077):
        want = torch.from_numpy(g[f"codes_{n}"].astype("int64"))
        got = M.mimi_encode(sd, _golden_waveform(n))[0]
        assert got.shape == want.shape == (32, M.encoded_frames(n))
        assert bool((got == want).all()), (n, int((got != want).sum()))
    assert len(set(g["codes_48077"][0].tolist())) > 10  # the quantizer is not stuck on one entry


@pytest.mark.parametrize("n", [1, 7, 1919, 1921, 24000 + 13])
def test_encode_restatement_matches_transformers_live(n):
    tr = pytest.importorskip("transformers")
    sd = _full_sd()
    m = tr.MimiModel(tr.MimiConfig(num_quantizers=32)).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    wav = torch.randn(1, 1, n, generator=torch.Generator().manual_seed(n)) * 0.3
    want = m.encode(wav, return_dict=True).audio_codes
    got = M.mimi_encode(sd, wav)
    assert got.shape == want.shape and bool((got == want).all())
    assert M.encoded_frames(n) == int(m.get_encoded_length(torch.tensor(n))) == got.shape[-1]

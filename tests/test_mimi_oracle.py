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


def test_decode_is_causal_prefix_exact():
    """Decoding a prefix gives the prefix of the decoded audio: what the streaming decoder relies on."""
    sd = M.synth_mimi_state_dict()
    g = torch.Generator().manual_seed(6)
    codes = torch.randint(0, 2048, (1, 32, 12), generator=g)
    full = M.mimi_decode(sd, codes)
    part = M.mimi_decode(sd, codes[:, :, :7])
    assert float((full[..., : 7 * 1920] - part).abs().max()) <= 1e-5

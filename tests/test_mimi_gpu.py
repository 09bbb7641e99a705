"""GPU parity of the CUDA Mimi decoder against the CPU oracle (which is pinned to transformers' MimiModel)."""
import numpy as np
import pytest
import torch

from oracle import mimi_oracle as M

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
_ENG = {}


def _engine():
    from sopro_b200.codec import MimiEngine

    if "e" not in _ENG:
        _ENG["sd"] = M.synth_mimi_state_dict()
        _ENG["e"] = MimiEngine(_ENG["sd"], 0, 32)
    return _ENG["e"], _ENG["sd"]


@pytest.mark.parametrize("B,T", [(1, 1), (2, 9), (1, 37), (3, 16)])
def test_decode_matches_oracle(B, T):
    """fp32 contraction order differs from the CPU's: tolerance 2e-4 of the waveform's peak."""
    eng, sd = _engine()
    codes = torch.randint(0, 2048, (B, 32, T), generator=torch.Generator().manual_seed(100 + T))
    want = M.mimi_decode(sd, codes)
    got = eng.decode(codes).cpu()
    assert got.shape == want.shape
    peak = float(want.abs().max())
    assert peak > 1e-3
    assert float((got - want).abs().max()) <= 2e-4 * max(1.0, peak)


def test_host_buffer_path_and_stream_decoder():
    from sopro_b200.codec import MimiCodec, MimiStreamDecoder

    eng, sd = _engine()
    codes = torch.randint(0, 2048, (1, 32, 20), generator=torch.Generator().manual_seed(3))
    full = eng.decode(codes).cpu().numpy()
    np.testing.assert_array_equal(eng.decode_host(codes.numpy()), full)
    codec = MimiCodec(32, device="cuda:0", state_dict=sd)
    dec = MimiStreamDecoder(codec)
    state, parts = None, []
    for a, b in [(0, 6), (6, 12), (12, 13), (13, 20)]:
        wav, state = dec.decode_step(codes[0, :, a:b].permute(1, 0), state)
        parts.append(wav.cpu().numpy())
    got = np.concatenate(parts, axis=1)
    assert got.shape == (1, 20 * 1920) and state.frames_seen == 20
    np.testing.assert_allclose(got[0], full[0, 0], rtol=0, atol=1e-5 * max(1.0, float(np.abs(full).max())))


def test_decode_full_signature():
    from sopro_b200.codec import MimiCodec

    _, sd = _engine()
    codec = MimiCodec(32, device="cuda:0", state_dict=sd)
    codes_tq = torch.randint(0, 2048, (5, 32), generator=torch.Generator().manual_seed(4))
    wav = codec.decode_full(codes_tq)
    assert wav.shape == (1, 1, 5 * 1920) and wav.dtype == torch.float32

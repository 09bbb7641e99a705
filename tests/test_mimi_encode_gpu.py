"""GPU parity of the CUDA Mimi ENCODER (waveform -> codes) against the CPU oracle, which reproduces transformers'
MimiModel.encode id for id (tests/test_mimi_oracle.py, tests/golden/mimi_encode.npz).  The codes are an argmin over
2048 distances per codebook: the pre-quantizer embeddings are held to 1e-4 of their peak, the ids to equality except
where the oracle's own two best distances are a near tie (then the rest of that frame's residual chain is not compared)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import mimi_oracle as M
from tests.golden.make_mimi_encode_golden import waveform

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
_ENG = {}


def _sd():
    if "sd" not in _ENG:
        sd = dict(M.synth_mimi_state_dict())
        sd.update(M.synth_mimi_encoder_state_dict())
        _ENG["sd"] = sd
    return _ENG["sd"]


def _engine():
    from sopro_b200.codec import MimiEncoderEngine

    if "e" not in _ENG:
        _ENG["e"] = MimiEncoderEngine(_sd(), 0, 32)
    return _ENG["e"]


def _compare_codes(sd, lat_want, got_qt, want_qt, rel=2e-3):
    """Every id equal, or the first differing codebook of a frame is a near tie in the oracle's own arithmetic (squared
    distances of the two candidates within `rel`).  Returns (frames with a tie, frames)."""
    Q, T = want_qt.shape
    ties = 0
    for t in range(T):
        diff = (got_qt[:, t] != want_qt[:, t]).nonzero()
        if diff.numel() == 0:
            continue
        q = int(diff[0])
        grp, i = ("semantic", q) if q < 1 else ("acoustic", q - 1)
        pre = f"quantizer.{grp}_residual_vector_quantizer."
        res = F.linear(lat_want[t], sd[pre + "input_proj.weight"].squeeze(-1))
        for j in range(i):  # the residual chain up to codebook q follows the (equal) earlier ids
            res = res - M.codebook(sd, pre + f"layers.{j}.codebook.")[want_qt[(0 if grp == "semantic" else 1) + j, t]]
        e = M.codebook(sd, pre + f"layers.{i}.codebook.")
        d_want = float((res - e[want_qt[q, t]]).pow(2).sum())
        d_got = float((res - e[got_qt[q, t]]).pow(2).sum())
        assert abs(d_got - d_want) <= rel * max(d_want, 1e-6), (t, q, d_got, d_want)
        ties += 1
    return ties, T


@pytest.mark.parametrize("n", [999, 5760, 13951, 48077])
def test_encode_matches_oracle_and_committed_transformers_codes(n):
    eng, sd = _engine(), _sd()
    wav = waveform(n)
    lat_want = M.mimi_encode_latent(sd, wav)[0]
    want = M.rvq_encode(sd, lat_want[None])[0]
    golden = torch.from_numpy(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mimi_encode.npz"))[f"codes_{n}"].astype("int64"))
    assert bool((want == golden).all())
    got, lat = eng.encode(wav, return_latent=True)
    got, lat = got.cpu(), lat.cpu()
    assert got.shape == want.shape == (32, eng.frames(n)) and eng.frames(n) == M.encoded_frames(n)
    peak = float(lat_want.abs().max())
    assert peak > 0.5
    assert float((lat - lat_want).abs().max()) <= 1e-4 * peak, (float((lat - lat_want).abs().max()), peak)
    ties, T = _compare_codes(sd, lat_want, got, want)
    print(f"n={n}: T={T}, latent err {float((lat - lat_want).abs().max()) / peak:.1e} of peak, ids equal "
          f"{float((got == want).float().mean()):.4f}, near-tie frames {ties}")
    assert ties <= max(1, T // 10)


@pytest.mark.parametrize("n", [1, 7, 1921])
def test_encode_tiny_inputs(n):
    """One sample, less than one conv kernel, one sample past a frame: the padding rules at every stride."""
    eng, sd = _engine(), _sd()
    wav = waveform(n)
    lat_want = M.mimi_encode_latent(sd, wav)[0]
    want = M.rvq_encode(sd, lat_want[None])[0]
    got, lat = eng.encode(wav, return_latent=True)
    assert got.shape == want.shape == (32, M.encoded_frames(n))
    assert float((lat.cpu() - lat_want).abs().max()) <= 1e-4 * max(1.0, float(lat_want.abs().max()))
    _compare_codes(sd, lat_want, got.cpu(), want)


def test_encode_long_input_past_the_attention_window():
    """12 s of audio = 300 transformer positions (> the 250-position window) and a ragged tail at every stride."""
    eng, sd = _engine(), _sd()
    n = 24000 * 12 + 1234
    wav = waveform(n)
    lat_want = M.mimi_encode_latent(sd, wav)[0]
    want = M.rvq_encode(sd, lat_want[None])[0]
    got, lat = eng.encode(wav, return_latent=True)
    got, lat = got.cpu(), lat.cpu()
    assert got.shape == want.shape
    peak = float(lat_want.abs().max())
    assert float((lat - lat_want).abs().max()) <= 1e-4 * peak
    ties, T = _compare_codes(sd, lat_want, got, want)
    assert T == 151 and ties <= T // 10, (ties, T)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    gw = wav.cuda()
    eng.encode(gw)
    ev[0].record()
    for _ in range(5):
        eng.encode(gw)
    ev[1].record()
    torch.cuda.synchronize()
    print(f"encode of {n / 24000:.1f} s of audio: {ev[0].elapsed_time(ev[1]) / 5:.2f} ms (fp32)")


def test_encode_is_deterministic_and_host_entry_point_agrees():
    eng = _engine()
    wav = waveform(13951)
    a = eng.encode(wav).cpu()
    b = eng.encode(wav.view(-1)).cpu()
    assert bool((a == b).all())
    h = torch.from_numpy(eng.encode_host(wav.view(-1).numpy()).astype("int64"))
    assert bool((a == h).all())
    assert int(a.min()) >= 0 and int(a.max()) < 2048


def test_encode_rejects_empty_input():
    from sopro_b200 import _lib

    eng = _engine()
    with pytest.raises(ValueError):
        eng.frames(0)
    with pytest.raises((ValueError, _lib.SoproError)):
        eng.encode(torch.zeros(0))


def test_codec_encode_file_roundtrip(tmp_path):
    """MimiCodec.encode_file (reference codec/mimi.py:41-63) on the CUDA encoder: PCM16 file -> trim / resample / crop on
    the host -> codes [T, 32]; decoding them gives T*1920 samples.  A decode-only state_dict refuses to encode."""
    from sopro_b200.audio import save_audio
    from sopro_b200.codec import MimiCodec

    codec = MimiCodec(32, device="cuda:0", state_dict=_sd(), precision="fp32")
    wav = waveform(24000 * 2)[0]
    path = str(tmp_path / "ref.wav")
    save_audio(path, wav, 24000)
    codes = codec.encode_file(path, crop_seconds=1.0)
    assert codes.dtype == torch.long and codes.shape[1] == 32 and 1 <= codes.shape[0] <= 13
    out = codec.decode_full(codes)
    assert out.shape == (1, 1, codes.shape[0] * 1920) and bool(torch.isfinite(out).all())
    # the model call alone equals the engine on the same samples
    assert bool((codec.encode_wav(wav) == _engine().encode(wav).permute(1, 0)).all())
    with pytest.raises(RuntimeError):
        MimiCodec(32, device="cuda:0", state_dict=M.synth_mimi_state_dict()).encode_wav(wav)

"""Host-side audio I/O used by ``encode_file`` / ``save_wav`` (reference audio.py; out of the hot path,
SURVEY.md §2 #14).  Reading/writing goes through soundfile or torchaudio when present and falls back to the
standard-library ``wave`` module (PCM16), so ``save_wav`` never needs an extra dependency."""
from __future__ import annotations

import os
import wave
from typing import Tuple

import numpy as np
import torch

from .config import TARGET_SR


def load_audio_file(path: str) -> Tuple[torch.Tensor, int]:
    """-> (mono float32 [1, T], sample rate); multi-channel files are averaged (reference audio.py:90-104)."""
    try:
        import soundfile as sf

        data, sr = sf.read(path, dtype="float32", always_2d=True)
        wav = torch.from_numpy(data).transpose(0, 1)
    except ImportError:
        try:
            import torchaudio

            wav, sr = torchaudio.load(path)
            wav = wav.float() / (2 ** 15) if wav.dtype == torch.int16 else wav.float()
        except ImportError:
            with wave.open(path, "rb") as f:
                sr, nch, width = f.getframerate(), f.getnchannels(), f.getsampwidth()
                if width != 2:
                    raise RuntimeError("wave fallback reads PCM16 only; install soundfile or torchaudio")
                pcm = np.frombuffer(f.readframes(f.getnframes()), dtype="<i2").reshape(-1, nch)
            wav = torch.from_numpy(pcm.astype(np.float32) / 32768.0).transpose(0, 1)
    if wav.size(0) > 1:
        wav = wav.mean(dim=0, keepdim=True)
    return wav.contiguous(), int(sr)


def resample(wav: torch.Tensor, sr_in: int, sr_out: int) -> torch.Tensor:
    if sr_in == sr_out:
        return wav
    import torchaudio.functional as AF  # same resampler as the reference (audio.py:107-117)

    return AF.resample(wav, sr_in, sr_out)


def trim_silence_energy(wav: torch.Tensor, sr: int, frame_ms: float = 25.0, hop_ms: float = 10.0, floor_db: float = -40.0,
                        pad_ms: float = 30.0, min_keep_sec: float = 0.5) -> torch.Tensor:
    """Energy VAD trim (reference audio.py:30-87): keep [first voiced frame - pad, last voiced frame + pad] where a
    frame is voiced if its energy is above max(peak - 40 dB, -40 dB); never trims below `min_keep_sec`."""
    one_d = wav.ndim == 1
    w = wav.unsqueeze(0) if one_d else wav
    T = w.shape[-1]
    flen, hop = max(1, int(sr * frame_ms / 1000.0)), max(1, int(sr * hop_ms / 1000.0))
    if T < int(sr * 0.1) or T < flen:
        return wav
    e_db = 10.0 * torch.log10(w.mean(dim=0, keepdim=True).unfold(-1, flen, hop).pow(2).mean(dim=-1).squeeze(0) + 1e-10)
    thr = max(float(e_db.max()) + floor_db, floor_db)
    idx = torch.nonzero(e_db > thr)
    if idx.numel() == 0:
        return wav
    pad = int(sr * pad_ms / 1000.0)
    start = max(0, int(idx[0, 0]) * hop - pad)
    end = min(T, int(idx[-1, 0]) * hop + flen + pad)
    if end - start < int(min_keep_sec * sr):
        return wav
    out = w[:, start:end]
    return out.squeeze(0) if one_d else out


def center_crop_audio(wav: torch.Tensor, win_samples: int) -> torch.Tensor:
    T = int(wav.shape[-1])
    if win_samples <= 0 or T <= win_samples:
        return wav
    s = (T - win_samples) // 2
    return wav[..., s: s + win_samples]


def save_audio(path: str, wav: torch.Tensor, sr: int = TARGET_SR) -> None:
    """Accepts [T], [C, T] or [B, C, T] (first item), writes mono (reference audio.py:120-144)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    w = wav.detach().float().cpu()
    if w.ndim == 3:
        w = w[0]
    if w.ndim == 1:
        w = w.unsqueeze(0)
    if w.ndim != 2:
        raise ValueError(f"Expected wav with 1-3 dims, got shape {tuple(wav.shape)}")
    mono = w.mean(dim=0).numpy()
    try:
        import soundfile as sf

        sf.write(path, mono, sr)
    except ImportError:
        pcm = np.clip(np.round(mono * 32767.0), -32768, 32767).astype("<i2")
        with wave.open(path, "wb") as f:
            f.setnchannels(1)
            f.setsampwidth(2)
            f.setframerate(int(sr))
            f.writeframes(pcm.tobytes())

"""Chunked streaming synthesis (reference streaming.py:12-152): every ``chunk_frames`` AR tokens the NAR refiner
runs over the new frames plus ``rf_nar`` frames of left context and the Mimi stream decoder emits their audio.
The AR kernel is launched ``chunk_frames`` frames at a time, so time-to-first-audio is prefill + one short
persistent launch + one NAR window + one Mimi decode."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List, Optional

import torch

from .codec import MimiDecodeState, MimiStreamDecoder
from .prefill import PreparedReference


@dataclass
class StreamConfig:
    chunk_frames: int = 16
    nar_context_frames: Optional[int] = None


class SoproTTSStreamer:
    def __init__(self, tts, cfg: Optional[StreamConfig] = None):
        self.tts = tts
        self.cfg = cfg or StreamConfig()
        self.mimi_stream = MimiStreamDecoder(tts.codec)

    @torch.inference_mode()
    def stream(self, text: str, *, ref_audio_path: Optional[str] = None, ref_tokens_tq: Optional[torch.Tensor] = None,
               ref: Optional[PreparedReference] = None, max_frames: int = 400, top_p: float = 0.9, temperature: float = 1.05,
               anti_loop: bool = True, style_strength: Optional[float] = None, ref_seconds: Optional[float] = None,
               chunk_frames: Optional[int] = None, nar_context_frames: Optional[int] = None,
               min_gen_frames: Optional[int] = None, seed: Optional[int] = None,
               generator: Optional[torch.Generator] = None) -> Iterator[torch.Tensor]:
        tts, model = self.tts, self.tts.model
        text_ids = tts.encode_text(text)
        if ref is None:
            ref = tts.prepare_reference(ref_audio_path=ref_audio_path, ref_tokens_tq=ref_tokens_tq, ref_seconds=ref_seconds)
        prep = model.prepare_conditioning(
            text_ids, ref, max_frames=max_frames,
            style_strength=float(style_strength if style_strength is not None else tts.cfg.style_strength))
        cf = int(chunk_frames if chunk_frames is not None else self.cfg.chunk_frames)
        ctx = nar_context_frames if nar_context_frames is not None else self.cfg.nar_context_frames
        ctx = int(model.rf_nar() if ctx is None else ctx)
        hist: List[int] = []
        emitted = 0
        state = MimiDecodeState()

        def refine_and_emit(end: int) -> Optional[torch.Tensor]:
            nonlocal emitted, state
            if end <= emitted:
                return None
            lo = max(0, emitted - ctx)
            toks = torch.as_tensor(hist[lo:end], device=tts.device, dtype=torch.long).unsqueeze(0)
            win = model.nar_refine(prep["cond_ar"][:, lo:end, :], toks).squeeze(0)
            wav, state = self.mimi_stream.decode_step(win[emitted - lo:, :], state)
            emitted = end
            return wav if wav.numel() > 0 else None

        for _t, tok, is_eos in model.ar_stream(prep, max_frames=max_frames, top_p=top_p, temperature=temperature,
                                               anti_loop=anti_loop, min_gen_frames=min_gen_frames, launch_frames=cf,
                                               seed=seed, generator=generator):
            if is_eos:  # streaming stops at the first EOS regardless of min_gen_frames (reference streaming.py:114-115)
                break
            hist.append(int(tok))
            if len(hist) % cf == 0:
                wav = refine_and_emit(len(hist))
                if wav is not None:
                    yield wav
        if emitted < len(hist):
            wav = refine_and_emit(len(hist))
            if wav is not None:
                yield wav


@torch.inference_mode()
def stream(tts, text: str, *, ref_audio_path: Optional[str] = None, ref_tokens_tq: Optional[torch.Tensor] = None,
           ref: Optional[PreparedReference] = None, chunk_frames: int = 6, **kwargs) -> Iterator[torch.Tensor]:
    streamer = SoproTTSStreamer(tts, StreamConfig(chunk_frames=chunk_frames))
    return streamer.stream(text, ref_audio_path=ref_audio_path, ref_tokens_tq=ref_tokens_tq, ref=ref,
                           chunk_frames=chunk_frames, **kwargs)

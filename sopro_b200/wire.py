"""The demo server's streaming wire format (reference demo/server.py:117-143, parsed by demo/static/app.js:860-900):

    b"SPRO" | u32 sample_rate | u32 channels |  then per chunk:  u32 byte_length | PCM16-LE samples

and its float -> PCM16 rule (clamp to [-1, 1], x 32767, truncate toward zero).  Host-side framing only: the samples come
from ``SoproTTS.stream`` (CUDA); the clamp / scale / int16 cast run on whatever device holds the chunk, so a streaming
server copies int16 (half the bytes) to the host."""
from __future__ import annotations

import struct
from typing import Iterable, Iterator, List, Tuple

import torch

MAGIC = b"SPRO"


def float_to_pcm16le(wav_1xt: torch.Tensor) -> bytes:
    """reference demo/server.py:117-124."""
    if wav_1xt.ndim == 1:
        wav_1xt = wav_1xt.unsqueeze(0)
    pcm = (wav_1xt.detach().clamp(-1.0, 1.0) * 32767.0).to(torch.int16)  # on the chunk's device
    return pcm.cpu().numpy().tobytes(order="C")


def stream_header(sr: int, channels: int = 1) -> bytes:
    """reference demo/server.py:138-140."""
    return MAGIC + struct.pack("<II", int(sr), int(channels))


def frame(payload: bytes) -> bytes:
    """reference demo/server.py:142-143."""
    return struct.pack("<I", len(payload)) + payload


def encode_stream(chunks: Iterable[torch.Tensor], sr: int = 24000, channels: int = 1) -> Iterator[bytes]:
    """Header, then one frame per audio chunk of ``SoproTTS.stream`` (what the demo's /v1/tts/stream endpoint sends)."""
    yield stream_header(sr, channels)
    for c in chunks:
        yield frame(float_to_pcm16le(c))


def parse_stream(data: bytes) -> Tuple[int, int, List[torch.Tensor]]:
    """Inverse of encode_stream (the browser client's parser, demo/static/app.js:860-900): -> (sr, channels, int16 chunks)."""
    if len(data) < 12 or data[:4] != MAGIC:
        raise ValueError("Bad stream header (magic mismatch).")
    sr, ch = struct.unpack("<II", data[4:12])
    out, pos = [], 12
    while pos < len(data):
        if pos + 4 > len(data):
            raise ValueError("truncated frame header")
        (n,) = struct.unpack("<I", data[pos:pos + 4])
        pos += 4
        if pos + n > len(data) or n % 2:
            raise ValueError("truncated frame payload")
        out.append(torch.frombuffer(bytearray(data[pos:pos + n]), dtype=torch.int16).clone() if n else torch.empty(0, dtype=torch.int16))
        pos += n
    return int(sr), int(ch), out

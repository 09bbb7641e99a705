"""The demo wire format (reference demo/server.py:117-143): known-answer bytes, the PCM16 rule's edge cases, round trip."""
import struct

import pytest
import torch

from sopro_b200 import wire


def test_header_and_frame_known_bytes():
    assert wire.stream_header(24000, 1) == b"SPRO" + bytes([0xC0, 0x5D, 0, 0]) + bytes([1, 0, 0, 0])
    assert wire.frame(b"\x01\x02") == bytes([2, 0, 0, 0, 1, 2])
    assert wire.frame(b"") == bytes(4)


def test_pcm16_rule_clamps_scales_and_truncates_toward_zero():
    x = torch.tensor([[0.0, 1.0, -1.0, 2.5, -7.0, 0.5, -0.5, 1e-5, 0.99999]])
    got = struct.unpack("<9h", wire.float_to_pcm16le(x))
    # clamp to [-1, 1], x 32767, int16 cast truncates toward zero (16383.5 -> 16383, -16383.5 -> -16383)
    assert got == (0, 32767, -32767, 32767, -32767, 16383, -16383, 0, 32766)
    assert wire.float_to_pcm16le(x[0]) == wire.float_to_pcm16le(x)  # 1-D input is one channel


def test_stream_round_trip_with_ragged_and_empty_chunks():
    g = torch.Generator().manual_seed(0)
    chunks = [torch.rand(1, n, generator=g) * 2 - 1 for n in (11520, 1, 0, 1920)]
    data = b"".join(wire.encode_stream(chunks, sr=24000))
    sr, ch, out = wire.parse_stream(data)
    assert (sr, ch) == (24000, 1) and [int(o.numel()) for o in out] == [11520, 1, 0, 1920]
    for c, o in zip(chunks, out):
        assert torch.equal(o, (c[0].clamp(-1, 1) * 32767.0).to(torch.int16))
    with pytest.raises(ValueError):
        wire.parse_stream(b"SPRX" + data[4:])
    with pytest.raises(ValueError):
        wire.parse_stream(data[:-1])

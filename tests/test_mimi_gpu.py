"""GPU parity of the CUDA Mimi decoder against the CPU oracle (which is pinned to transformers' MimiModel)."""
import numpy as np
import pytest
import torch

from oracle import mimi_oracle as M

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
_ENG = {}


def _engine(precision="fp32"):
    from sopro_b200.codec import MimiEngine

    if "e" not in _ENG:
        _ENG["sd"] = M.synth_mimi_state_dict()
        _ENG["e"] = MimiEngine(_ENG["sd"], 0, 32)
    _ENG["e"].set_precision(precision)
    return _ENG["e"], _ENG["sd"]


@pytest.mark.parametrize("B,T", [(1, 1), (2, 9), (1, 37), (3, 16)])
def test_decode_matches_oracle(B, T):
    """fp32 mode; the contraction order differs from the CPU's: tolerance 2e-4 of the waveform's peak."""
    eng, sd = _engine("fp32")
    codes = torch.randint(0, 2048, (B, 32, T), generator=torch.Generator().manual_seed(100 + T))
    want = M.mimi_decode(sd, codes)
    got = eng.decode(codes).cpu()
    assert got.shape == want.shape
    peak = float(want.abs().max())
    assert peak > 1e-3
    assert float((got - want).abs().max()) <= 2e-4 * max(1.0, peak)


@pytest.mark.parametrize("B,T", [(1, 1), (2, 9), (1, 37), (3, 16), (1, 150), (2, 203)])
def test_decode_tensor_core_mode(B, T):
    """Default mode: bf16 operands on the tcgen05 tensor cores, fp32 accumulation.  Stated tolerance: max error
    2e-2 of the waveform's peak and relative RMS error 1e-2 against the fp32 oracle."""
    eng, sd = _engine("bf16_tc")
    codes = torch.randint(0, 2048, (B, 32, T), generator=torch.Generator().manual_seed(100 + T))
    want = M.mimi_decode(sd, codes)
    got = eng.decode(codes).cpu()
    assert got.shape == want.shape and bool(torch.isfinite(got).all())
    peak = float(want.abs().max())
    err = got - want
    assert float(err.abs().max()) <= 2e-2 * peak, (float(err.abs().max()), peak)
    assert float(err.pow(2).mean().sqrt()) <= 1e-2 * float(want.pow(2).mean().sqrt())
    # against the bf16-operand model of this mode (oracle/mimi_oracle.py) the distance should be far smaller (accumulation
    # order + rare bf16 rounding flips); reported here, to be tightened into the bound once it has GPU history
    emu = M.mimi_decode_bf16_operands(sd, codes)
    d = float((got - emu).abs().max())
    print(f"tensor-core mode B={B} T={T}: max err vs fp32 oracle {float(err.abs().max()) / peak:.2e} of peak, "
          f"vs bf16-operand model {d / peak:.2e} of peak")
    assert d <= 2.6e-2 * peak


def test_full_size_decode_properties():
    """BASELINE.json's standalone configuration (10k frames) through size-independent properties: the decoder is
    causal, so the first 300 frames of the 10k-frame waveform equal a 300-frame decode bit for bit (same kernels, other
    grid sizes); the tensor-core result stays within the stated tolerance of the fp32 mode at full size; a batch of 25
    x 400 frames equals the same utterances decoded one by one."""
    eng, _ = _engine("bf16_tc")
    codes = torch.randint(0, 2048, (1, 32, 10000), generator=torch.Generator().manual_seed(5))
    big = eng.decode(codes)
    assert big.shape == (1, 1, 10000 * 1920) and bool(torch.isfinite(big).all())
    small = eng.decode(codes[:, :, :300])
    assert torch.equal(big[..., : 300 * 1920], small)
    eng.set_precision("fp32")
    ref32 = eng.decode(codes[:, :, :2000])
    eng.set_precision("bf16_tc")
    err = (big[..., : 2000 * 1920] - ref32).abs().max()
    assert float(err) <= 2e-2 * float(ref32.abs().max())
    # against the CPU oracle directly: the decoder is causal, so the oracle's decode of the first 700 frames is the
    # reference for frames [0, 700) of the 10k-frame waveform: the start [0, 300) and a mid-stream slice [400, 700)
    # (past the 125-frame attention window) in both modes
    want = M.mimi_decode(_ENG["sd"], codes[:, :, :700])
    peak = float(want.abs().max())
    for lo, hi in ((0, 300), (400, 700)):
        sl = slice(lo * 1920, hi * 1920)
        e_tc = float((big[..., sl].cpu() - want[..., sl]).abs().max())
        e_32 = float((ref32[..., sl].cpu() - want[..., sl]).abs().max())
        print(f"10k-frame decode vs oracle, frames [{lo},{hi}): tensor-core {e_tc / peak:.2e} of peak, fp32 {e_32 / peak:.2e} of peak")
        assert e_tc <= 2e-2 * peak, (lo, hi, e_tc, peak)
        assert e_32 <= 2e-4 * max(1.0, peak), (lo, hi, e_32, peak)
    batch = codes.view(1, 32, 25, 400).permute(2, 1, 0, 3).reshape(25, 32, 400).contiguous()
    wb = eng.decode(batch)
    for i in (0, 11, 24):
        assert torch.equal(wb[i: i + 1], eng.decode(batch[i: i + 1]))


def _im2col(x, taps, dil, pad):
    B, R, Cin = x.shape
    cols = []
    for j in range(taps):
        sh = j * dil - pad
        y = torch.zeros_like(x)
        lo, hi = max(0, -sh), min(R, R - sh)
        if hi > lo:
            y[:, lo:hi] = x[:, lo + sh:hi + sh]
        cols.append(y)
    return torch.cat(cols, dim=-1)


TC_GEMM_CASES = [
    # B, rows, cin, taps, dil, pad, N, bias_mod, epi, out_elu
    (2, 300, 512, 1, 1, 0, 1536, 0, 0, 0),      # QKV
    (1, 129, 512, 7, 1, 6, 1024, 1024, 0, 1),   # conv0 -> ELU'd bf16
    (2, 50, 1024, 2, 1, 1, 4096, 512, 0, 1),    # ConvTranspose stride 8 as a 2-tap conv
    (1, 1000, 256, 1, 1, 0, 512, 512, 3, 1),    # res conv k=1 + skip
    (3, 77, 64, 3, 1, 2, 32, 32, 0, 0),         # narrowest layer (N=32, K=192)
    (1, 260, 2048, 1, 1, 0, 512, 0, 2, 0),      # fc2 + LayerScale residual
    (2, 5, 512, 1, 1, 0, 2048, 0, 1, 0),        # fc1 + GELU, fewer rows than one tile
    (1, 200, 128, 3, 2, 4, 640, 128, 0, 0),     # dilation 2, N = 10 x 64
    (1, 500, 32, 1, 1, 0, 64, 64, 3, 1),        # last ResnetBlock's k=1 conv: 32 channels -> 64-byte rows, K = 32
    (2, 100, 32, 3, 1, 2, 128, 0, 0, 0),        # 32-channel input with taps
]


@pytest.mark.parametrize("case", TC_GEMM_CASES, ids=lambda c: "x".join(map(str, c)))
def test_tc_gemm_matches_torch(case):
    """The tcgen05 implicit GEMM alone against torch fp32 on the same bf16-rounded operands (differences are
    accumulation order only): 1e-3 of the output scale for fp32 results, one bf16 ulp (2^-8 relative) for bf16."""
    import ctypes as C

    from sopro_b200 import _lib

    lib = _lib.load()
    B, R, cin, taps, dil, pad, N, bias_mod, epi, out_elu = case
    g = torch.Generator().manual_seed(sum(case))
    dev = torch.device("cuda:0")
    x = (torch.randn(B, R, cin, generator=g)).to(torch.bfloat16).to(dev)
    w = (torch.randn(N, taps * cin, generator=g) / (taps * cin) ** 0.5).to(torch.bfloat16).to(dev)
    bias = torch.randn(bias_mod, generator=g).to(dev) if bias_mod else None
    res = torch.randn(B, R, N, generator=g).to(dev) if epi in (2, 3) else None
    scale = torch.rand(N, generator=g).to(dev) if epi == 2 else None
    acc = _im2col(x.float(), taps, dil, pad) @ w.float().t()
    if bias is not None:
        acc = acc + bias.repeat(N // bias_mod)
    if epi == 1:
        acc = torch.nn.functional.gelu(acc)
    elif epi == 2:
        acc = res + scale * acc
    elif epi == 3:
        acc = res + acc
    of = torch.full((B, R, N), float("nan"), device=dev)
    oh = torch.full((B, R, N), float("nan"), device=dev, dtype=torch.bfloat16)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
    _lib.check(lib.sopro_debug_tc_gemm(p(x), B, R, cin, taps, dil, pad, p(w), N, p(bias), bias_mod, epi, p(res), p(scale),
                                       p(of), p(oh), out_elu, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    s = float(acc.abs().max())
    assert float((of - acc).abs().max()) <= 1e-3 * s
    want_h = torch.nn.functional.elu(acc) if out_elu else acc
    assert float((oh.float() - want_h).abs().max()) <= 2 ** -8 * s + 1e-3 * s


@pytest.mark.parametrize("precision", ["fp32", "bf16_tc"])
def test_host_buffer_path_and_stream_decoder(precision):
    from sopro_b200.codec import MimiCodec, MimiStreamDecoder

    eng, sd = _engine(precision)
    codes = torch.randint(0, 2048, (1, 32, 20), generator=torch.Generator().manual_seed(3))
    full = eng.decode(codes).cpu().numpy()
    np.testing.assert_array_equal(eng.decode_host(codes.numpy()), full)
    codec = MimiCodec(32, device="cuda:0", state_dict=sd, precision=precision)
    dec = MimiStreamDecoder(codec)
    state, parts = None, []
    for a, b in [(0, 6), (6, 12), (12, 13), (13, 20)]:
        wav, state = dec.decode_step(codes[0, :, a:b].permute(1, 0), state)
        parts.append(wav.cpu().numpy())
    got = np.concatenate(parts, axis=1)
    assert got.shape == (1, 20 * 1920) and state.frames_seen == 20
    if precision == "fp32":  # the persistent-state stream decoder computes every sample in the one-shot decode's order
        np.testing.assert_array_equal(got[0], full[0, 0])
    else:  # tensor-core mode: same tcgen05 tiles, the attention core runs in fp32 over the K/V ring (tolerance 1e-2 of peak)
        np.testing.assert_allclose(got[0], full[0, 0], rtol=0, atol=1e-2 * float(np.abs(full).max()))


def test_small_decodes_replay_from_graphs_identically():
    """<= 64 frames go through the library's CUDA-graph cache: first call captures, later calls replay; both equal the
    plain launches bit for bit, also after a large decode reallocated the workspace (graphs are dropped then)."""
    eng, _ = _engine("bf16_tc")
    codes = torch.randint(0, 2048, (2, 32, 7), generator=torch.Generator().manual_seed(9))
    other = torch.randint(0, 2048, (2, 32, 7), generator=torch.Generator().manual_seed(10))
    eng.set_graphs(False)
    want, want_other = eng.decode(codes).cpu(), eng.decode(other).cpu()
    eng.set_graphs(True)
    assert torch.equal(eng.decode(codes).cpu(), want)        # capture + first replay
    assert torch.equal(eng.decode(other).cpu(), want_other)  # replay with new inputs
    eng.decode(torch.randint(0, 2048, (4, 32, 120)))         # grows the workspace
    assert torch.equal(eng.decode(codes).cpu(), want)


def test_decode_full_signature():
    from sopro_b200.codec import MimiCodec

    _, sd = _engine()
    codec = MimiCodec(32, device="cuda:0", state_dict=sd)
    codes_tq = torch.randint(0, 2048, (5, 32), generator=torch.Generator().manual_seed(4))
    wav = codec.decode_full(codes_tq)
    assert wav.shape == (1, 1, 5 * 1920) and wav.dtype == torch.float32


@pytest.mark.parametrize("mode", ["fp32", "bf16_tc"])
def test_stream_decode_step_equals_the_full_decode(mode):
    """sopro_mimi_decode_step carries the K/V rings and every conv's left context: chunks of ragged sizes (1 frame,
    the default 6, 16, a 40-frame chunk that is split internally, ...) over 310 frames -- 620 transformer positions,
    far past the 250-position window and past the ring's wrap-around -- concatenate to the one-shot decode.  fp32
    mode: bit for bit (every output element is computed in the same order).  Tensor-core mode: the dense blocks are the
    same tcgen05 tiles, only the attention core runs in fp32 on the ring; the stated tolerance is 1e-2 of the peak vs the
    one-shot tensor-core decode and the mode's 2e-2 vs the fp32 oracle."""
    eng, sd = _engine(mode)
    T = 310
    codes = torch.randint(0, 2048, (1, 32, T), generator=torch.Generator().manual_seed(77))
    full = eng.decode(codes).reshape(1, -1)
    st = eng.stream(16)
    sizes, pos, parts = [1, 6, 6, 16, 3, 40, 6, 2, 64, 6], 0, []
    i = 0
    while pos < T:
        n = min(sizes[i % len(sizes)], T - pos)
        parts.append(st.step(codes[0, :, pos:pos + n]))
        pos += n
        i += 1
    assert st.frames == T
    got = torch.cat(parts, dim=1)
    assert got.shape == full.shape
    peak = float(full.abs().max())
    err = float((got - full).abs().max())
    print(f"stream vs one-shot [{mode}]: max err {err / peak:.2e} of peak")
    if mode == "fp32":
        assert torch.equal(got, full)
    else:
        assert err <= 1e-2 * peak, (err, peak)
        want = M.mimi_decode(sd, codes[:, :, :150]).reshape(1, -1)
        assert float((got[:, : 150 * 1920].cpu() - want).abs().max()) <= 2e-2 * float(want.abs().max())
    # reset -> the same stream object decodes a new utterance from frame 0; host-buffer entry point
    st.reset()
    assert st.frames == 0
    again = st.step_host(codes[0, :, :9].numpy())
    one = eng.decode(codes[:, :, :9]).reshape(1, -1).cpu().numpy()
    if mode == "fp32":
        assert np.array_equal(again, one)
    else:
        assert float(np.abs(again - one).max()) <= 1e-2 * peak


def test_two_streams_are_independent_and_state_is_per_stream():
    eng, _ = _engine("fp32")
    a = torch.randint(0, 2048, (32, 30), generator=torch.Generator().manual_seed(1))
    b = torch.randint(0, 2048, (32, 30), generator=torch.Generator().manual_seed(2))
    sa, sb = eng.stream(8), eng.stream(8)
    outa, outb = [], []
    for lo in range(0, 30, 6):  # interleaved
        outa.append(sa.step(a[:, lo:lo + 6]))
        outb.append(sb.step(b[:, lo:lo + 6]))
    assert torch.equal(torch.cat(outa, 1), eng.decode(a.unsqueeze(0)).reshape(1, -1))
    assert torch.equal(torch.cat(outb, 1), eng.decode(b.unsqueeze(0)).reshape(1, -1))


def test_out_of_range_codes_are_reported_not_read():
    """An uncut EOS id (2048) must not index past the codebook: the Python layer raises IndexError like the reference's
    embedding lookup, the host C-ABI path returns an error, and the device path clamps + flags (sopro_mimi_check)."""
    import ctypes as C

    from sopro_b200 import _lib

    eng, _ = _engine("fp32")
    bad = torch.randint(0, 2048, (1, 32, 4))
    bad[0, 3, 2] = 2048
    with pytest.raises(IndexError):
        eng.decode(bad)
    with pytest.raises(_lib.SoproError):
        eng.decode_host(bad.numpy())
    dev = bad.to("cuda:0", torch.int32).contiguous()
    wav = torch.empty(1, 1, 4 * 1920, device="cuda:0")
    _lib.check(eng.lib.sopro_mimi_decode(eng._h, dev.data_ptr(), 1, 4, wav.data_ptr(), int(torch.cuda.current_stream().cuda_stream)))
    with pytest.raises(_lib.SoproError):
        eng.check()
    eng.check()  # the flag is cleared by the failed check
    assert bool(torch.isfinite(wav).all())

"""Host-side control flow of the public API (synthesize / synthesize_batch / stream) WITHOUT a GPU.

The product has no CPU path, so the four CUDA engines (prefill, AR, NAR, Mimi) are replaced here, in the test only, by
fakes that answer through the CPU oracles (oracle/ar_oracle.py, oracle/nar_oracle.py, oracle/mimi_oracle.py and the torch
restatement of the prefill).  What is under test is everything around the kernels in sopro_b200/model.py, streaming.py and
codec.py: lazy noise blocks, launch chunking, EOS handling, the ragged NAR batch, right-padding and cutting of the batched
Mimi decode, streaming chunk sizes, session check-out.
"""
import numpy as np
import pytest
import torch

from oracle import ar_oracle as O
from oracle import mimi_oracle as M
from oracle import nar_oracle as N
from sopro_b200 import prefill as P
from sopro_b200.codec import MimiStreamDecoder
from sopro_b200.config import SoproTTSConfig
from sopro_b200.model import SoproModel, SoproTTS
from sopro_b200.tokenizer import IdsTokenizer
from sopro_b200.weights import synth_state_dict

torch.set_grad_enabled(False)


class _FakeSession:
    """Same contract as sopro_b200.engine.ArSession (begin / run / read / position), answered by the oracle."""

    def __init__(self, owner, B, steps, L):
        self.o, self.maxB, self.max_steps = owner, B, steps

    def begin(self, cond, txt, lens, noise, samp):
        samps = [samp] * cond.shape[0] if not isinstance(samp, (list, tuple)) else list(samp)
        self.B, self.steps = int(cond.shape[0]), int(cond.shape[1])
        self.noise = noise  # kept by reference: ar_stream() fills rows just before each run()
        self.V = self.o.cfg.ar_vocab()
        self.toks = np.zeros((self.B, self.steps), dtype=np.int32)
        self.n = np.zeros(self.B, dtype=np.int32)
        self.done = np.zeros(self.B, dtype=np.int32)
        self.t_pos = 0
        self.full = [torch.ones(self.steps, self.V) for _ in range(self.B)]
        self.gens, self.samps = [], samps
        for b in range(self.B):
            L = int(lens[b])
            s = samps[b]
            osamp = O.ArSampling(top_p=s.top_p, temperature=s.temperature, anti_loop=bool(s.anti_loop), loop_streak=s.loop_streak,
                                 recovery_top_p=s.recovery_top_p, recovery_temp=s.recovery_temp, min_gen_frames=s.min_gen_frames,
                                 top_k=s.top_k, repetition_penalty=s.repetition_penalty)
            self.gens.append(O.ar_stream(self.o.sd, self.o.cfg, cond[b:b + 1].cpu(), txt[b:b + 1, :L].cpu(),
                                         torch.ones(1, L, dtype=torch.bool), max_frames=self.steps - 1, sampling=osamp,
                                         noise_tv=self.full[b]))

    def run(self, n_steps=None):
        end = min(self.steps, self.t_pos + (int(n_steps) if n_steps is not None else self.steps))
        for b in range(self.B):
            k = int(self.noise.shape[2])
            while not self.done[b] and self.n[b] < end:
                t = int(self.n[b])
                self.full[b][t, :k] = self.noise[b, t].cpu()
                try:
                    _t, tok, is_eos = next(self.gens[b])
                except StopIteration:
                    self.done[b] = 1
                    break
                self.toks[b, t] = tok
                self.n[b] = t + 1
                s = self.samps[b]
                if is_eos and (s.stop_on_first_eos or t + 1 >= s.min_gen_frames):
                    self.done[b] = 1
                if t + 1 >= self.steps:
                    self.done[b] = 1
        self.t_pos = end

    def read(self):
        return self.toks.copy(), self.n.copy(), self.done.copy()

    @property
    def position(self):
        return self.t_pos

    def close(self):
        pass


class _FakeArEngine:
    def __init__(self, cfg, sd):
        self.cfg, self.sd = cfg, sd

    def session(self, B, steps, L):
        return _FakeSession(self, B, steps, L)


class _FakePrefill:
    """sopro_b200.prefill_cuda.PrefillEngine.run through the torch restatement (bit-equal to the reference on CPU)."""

    def __init__(self, m):
        self.m = m

    def run(self, text_ids, ref, *, n_frames, style_strength):
        preps = [P.prepare_conditioning(self.m.sd, self.m.cfg, ids, ref, max_frames=n_frames - 1, device="cpu",
                                        style_strength=style_strength, text_pos=self.m.text_pos, frame_pos=self.m.frame_pos)
                 for ids in text_ids]
        lens = [int(p["txt_seq"].size(1)) for p in preps]
        txt = torch.zeros(len(preps), max(lens), int(self.m.cfg.d_model))
        for i, p in enumerate(preps):
            txt[i, : lens[i]] = p["txt_seq"][0]
        return txt, lens, torch.cat([p["txt_pool"] for p in preps]), torch.cat([p["cond_ar"] for p in preps])


class _FakeRefPrep:
    """sopro_b200.prefill_cuda.RefPrepEngine.run through the torch restatement."""

    def __init__(self, m):
        self.m = m

    def run(self, ref_tokens_tq):
        r = P.prepare_reference(self.m.sd, self.m.cfg, ref_tokens_tq, torch.device("cpu"))
        return r.sv_ref, r.ref_seq, r.ref_kv_caches


class _FakeNar:
    """sopro_b200.nar.NarEngine.refine through oracle/nar_oracle.py, utterance by utterance over its valid frames."""

    def __init__(self, m):
        self.m = m

    def refine(self, cond, rvq1, lens=None):
        B, T, _ = cond.shape
        out = torch.zeros(B, T, int(self.m.cfg.num_codebooks), dtype=torch.long)
        for b in range(B):
            n = T if lens is None else int(lens[b])
            if n:
                out[b, :n] = N.nar_refine(self.m.sd, self.m.cfg, cond[b:b + 1, :n].cpu(), rvq1[b:b + 1, :n].cpu().long())[0][0]
        return out


class _FakeMimiStream:
    def __init__(self, eng):
        self.eng, self.hist = eng, None

    def reset(self):  # a pooled state handed to the next utterance (MimiStream.reset)
        self.hist = None

    def step(self, codes_qn, trusted=False):
        self.hist = codes_qn if self.hist is None else torch.cat([self.hist, codes_qn], dim=1)
        wav = self.eng.decode(self.hist.unsqueeze(0)).reshape(1, -1)
        return wav[:, (self.hist.shape[1] - codes_qn.shape[1]) * 1920:]


class _FakeMimiEngine:
    hop = 1920
    precision = "oracle"

    def __init__(self, msd):
        self.msd, self.calls = msd, []

    def stream(self, max_chunk_frames=16):
        return _FakeMimiStream(self)

    def decode(self, codes_bqt):
        self.calls.append(tuple(codes_bqt.shape))
        if codes_bqt.shape[2] == 0:  # like MimiEngine.decode
            return torch.zeros(codes_bqt.shape[0], 1, 0)
        return M.mimi_decode(self.msd, codes_bqt.long().cpu())


class _FakeCodec:
    def __init__(self, msd):
        self.engine, self.device = _FakeMimiEngine(msd), torch.device("cpu")

    def decode_full(self, codes_tq):
        return self.engine.decode(codes_tq.permute(1, 0).unsqueeze(0))


@pytest.fixture(scope="module")
def tts():
    cfg = SoproTTSConfig()
    sd = synth_state_dict(cfg, text_vocab=1000, seed=0)
    sd["ar.head.bias"] = sd["ar.head.bias"].clone()
    sd["ar.head.bias"][int(cfg.codebook_size)] += 2.5  # EOS a few times more likely than a code: ragged lengths
    m = object.__new__(SoproModel)  # the real constructor insists on a CUDA device and builds the CUDA engine
    m.cfg, m.device, m.eos_id, m.weight_dtype = cfg, torch.device("cpu"), int(cfg.codebook_size), "fp32"
    m.engine = _FakeArEngine(cfg, sd)
    skip = ("ar.blocks.", "ar.head.", "ar.norm.")
    m.sd = {k: v.float() for k, v in sd.items() if not k.startswith(skip) and v.is_floating_point()}
    m.text_pos = P.sinusoid_table(int(cfg.max_text_len) + 8, int(cfg.d_model), "cpu")
    m.frame_pos = P.sinusoid_table(int(cfg.pos_emb_max) + 8, int(cfg.d_model), "cpu")
    import threading

    m._sessions, m._sessions_busy, m._sessions_lock = {}, set(), threading.Lock()
    m.prefill, m.nar, m.refprep = _FakePrefill(m), _FakeNar(m), _FakeRefPrep(m)
    t = SoproTTS(model=m, cfg=cfg, tokenizer=IdsTokenizer(1000), codec=_FakeCodec(M.synth_mimi_state_dict()), device="cpu")
    t.ref = t.prepare_reference(ref_tokens_tq=torch.randint(0, 2048, (12, 32), generator=torch.Generator().manual_seed(7)))
    return t


TEXTS = ["3 14 15 92 65 35", " ".join(str(7 * i + 1) for i in range(15)), "8 9", "27 18 28 18"]
SEEDS = [1, 2, 3, 4]
KW = dict(max_frames=20, min_gen_frames=3)


def test_batch_equals_single_with_ragged_lengths(tts):
    wavs = tts.synthesize_batch(TEXTS, ref=tts.ref, seeds=SEEDS, **KW)
    lens = [w.shape[-1] // 1920 for w in wavs]
    assert len(set(lens)) >= 2, f"the case must produce different lengths to exercise padding, got {lens}"
    assert all(w.shape[:2] == (1, 1) and w.shape[-1] % 1920 == 0 for w in wavs)
    for text, seed, w in zip(TEXTS, SEEDS, wavs):
        single = tts.synthesize(text, ref=tts.ref, seed=seed, **KW)
        assert single.shape == w.shape, (single.shape, w.shape)
        np.testing.assert_allclose(w.numpy(), single.numpy(), rtol=0, atol=1e-5)
    # the batch went through a padded decode: at least one call carried more than one utterance
    assert any(c[0] > 1 for c in tts.codec.engine.calls)


def test_tokens_stop_at_first_eos_and_global_rng_is_settled(tts):
    ids = tts.encode_text(TEXTS[1])
    torch.manual_seed(11)
    a = tts.model.generate_tokens(ids, tts.ref, style_strength=1.0, **KW)
    state = torch.get_rng_state()
    T = a.shape[0]
    assert a.shape[1] == 32 and 0 < T <= KW["max_frames"] + 1 and int(a[:, 0].max()) < 2048  # cut before the EOS
    # the global generator sits exactly after the rows the reference would have drawn: T code frames + the EOS step
    torch.manual_seed(11)
    drawn = T + 1 if T < KW["max_frames"] + 1 else T
    torch.empty(drawn, tts.cfg.ar_vocab()).exponential_(1.0)
    assert torch.equal(torch.get_rng_state(), state)
    b = tts.model.generate_tokens(ids, tts.ref, style_strength=1.0, seed=11, **KW)
    assert torch.equal(a, b)


def test_stream_chunks_cover_the_utterance(tts):
    full = tts.synthesize(TEXTS[1], ref=tts.ref, seed=2, **KW)
    T = full.shape[-1] // 1920
    chunks = list(tts.stream(TEXTS[1], ref=tts.ref, seed=2, chunk_frames=4, **KW))
    want = [4] * (T // 4) + ([T % 4] if T % 4 else [])
    assert [c.shape for c in chunks] == [(1, k * 1920) for k in want]
    assert all(bool(torch.isfinite(c).all()) for c in chunks)


def test_stream_decoder_is_prefix_exact_through_the_codec_interface(tts):
    codes = torch.randint(0, 2048, (9, 32), generator=torch.Generator().manual_seed(5))
    full = tts.codec.decode_full(codes)
    dec, state, parts = MimiStreamDecoder(tts.codec), None, []
    for a, b in [(0, 4), (4, 5), (5, 9)]:
        w, state = dec.decode_step(codes[a:b], state)
        parts.append(w)
    np.testing.assert_allclose(torch.cat(parts, dim=1).numpy(), full.reshape(1, -1).numpy(), rtol=0, atol=1e-6)
    assert state.frames_seen == 9 and state.samples_emitted == 9 * 1920


def test_stream_nar_windows_follow_the_reference(tts, monkeypatch):
    """reference streaming.py:80-104: every chunk_frames tokens the NAR refiner sees the new frames plus
    nar_context_frames of left context, and only the new frames' codes go to the Mimi stream decoder."""
    seen = []
    real = tts.model.nar_refine

    def spy(cond, rvq1, lens=None):
        seen.append((int(cond.shape[1]), rvq1[0].tolist()))
        return real(cond, rvq1, lens)

    monkeypatch.setattr(tts.model, "nar_refine", spy)
    prep = tts.model.prepare_conditioning(tts.encode_text(TEXTS[1]), tts.ref, max_frames=KW["max_frames"],
                                          style_strength=tts.cfg.style_strength)
    toks = []
    for _t, tok, is_eos in tts.model.ar_stream(prep, seed=2, **KW):
        if is_eos:
            break
        toks.append(tok)
    chunks = list(tts.stream(TEXTS[1], ref=tts.ref, seed=2, chunk_frames=4, nar_context_frames=3, **KW))
    T, cf, ctx = len(toks), 4, 3
    ends = list(range(cf, T + 1, cf)) + ([T] if T % cf else [])
    want, emitted = [], 0
    for e in ends:
        lo = max(0, emitted - ctx)
        want.append((e - lo, toks[lo:e]))
        emitted = e
    assert seen == want
    assert sum(c.shape[1] for c in chunks) == T * 1920


def test_baseline_config0_plumbing(tts):
    """BASELINE.json configs[0]: batch 1, 50-token sentence (52 ids with BOS/EOS), 3 s reference (38 frames) — the
    reference's CPU-runnable case, here through the host pipeline with the oracle-backed engines."""
    ref = tts.prepare_reference(ref_tokens_tq=torch.randint(0, 2048, (38, 32), generator=torch.Generator().manual_seed(7)))
    assert ref.ref_tokens_btq.shape == (1, 38, 32) and ref.sv_ref.shape == (1, 192) and ref.ref_seq.shape == (1, 38, 384)
    text = " ".join(str(17 * i + 5) for i in range(50))
    assert tts.encode_text(text).numel() == 52
    # like the reference's generate_tokens (model.py:371-384) synthesis ends at the FIRST EOS whatever min_gen_frames
    # says; this fixture makes EOS likely, so take the first seed that yields a few frames
    ids = tts.encode_text(text)
    seed, T = next((s_, t_) for s_ in range(1, 30)
                   for t_ in [tts.model.generate_tokens(ids, ref, max_frames=12, style_strength=tts.cfg.style_strength, seed=s_).shape[0]]
                   if t_ >= 3)
    wav = tts.synthesize(text, ref=ref, max_frames=12, seed=seed)
    assert wav.shape == (1, 1, T * 1920) and wav.dtype == torch.float32 and bool(torch.isfinite(wav).all())
    assert torch.equal(wav, tts.synthesize(text, ref=ref, max_frames=12, seed=seed))

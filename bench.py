#!/usr/bin/env python
"""bench.py — AR frames/s of the Sopro hot path on N B200s (one process per GPU).

Workload (BASELINE.json configs[2]; x N GPUs it is configs[3]): per GPU a batch of 64
independent 400-frame utterances (401 AR steps: model.py:242), 52 text tokens each, bf16
weight storage, fp32 arithmetic, default sampler, EOS never terminates.  A bench "step" is one
full pass of the hot path over that batch: text-K/V build + ONE persistent AR kernel launch
(64 x 401 frames).  Data parallel, no data-path collective ("weak" scaling); NCCL is used once,
to broadcast the weights from rank 0.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the reference algorithm on the host CPU (oracle port)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

BATCH_PER_GPU = 64
FRAMES = 400
STEPS_AR = FRAMES + 1
TEXT_LEN = 52
S_UTT_BYTES = 3280  # cond row + embedding row + noise + token per utterance-step
WORKLOAD = "batch=64/GPU non-streaming, 400-frame utterances (401 AR steps), L_text=52, bf16 weights, fp32 math"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _inputs(cfg, rank, B, steps, L):
    from sopro_b200.sampling import noise_tape
    from sopro_b200.weights import hash_uniform

    D = int(cfg.d_model)
    s3 = np.float32(np.sqrt(3.0))
    cond = torch.from_numpy(hash_uniform(B * steps * D, 7_000_000 + rank) * s3).view(B, steps, D)
    txt = torch.from_numpy(hash_uniform(B * L * D, 8_000_000 + rank) * s3).view(B, L, D)
    noise = torch.stack([noise_tape(steps, cfg.ar_vocab(), seed=1234 + rank * B + i, keep=50) for i in range(B)])
    return cond.contiguous(), txt.contiguous(), noise.contiguous()


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason samples during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": int(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def usable_cpus() -> int:
    """Host threads this process may really use: affinity mask and cgroup quota, not the box's core count."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            q, per = f.read().split()
            if q != "max":
                n = min(n, max(1, int(float(q) / float(per))))
    except Exception:
        pass
    return max(1, n)


def pick_threads(cfg, sd):
    """The reference's AR step is ~1,240 tiny ATen calls: more intra-op threads than a
    handful only add fork/join cost.  Try a few counts on a 24-frame probe and keep the fastest, so the CPU
    arm gets the best configuration this host offers; the chosen count is reported as `cores`."""
    from oracle import ar_oracle as O

    top = usable_cpus()
    cands = sorted({c for c in (1, 4, 8, 16, 32, top) if c <= top})
    cond, txt, _ = _inputs(cfg, 0, 1, 25, TEXT_LEN)
    mask = torch.ones(1, TEXT_LEN, dtype=torch.bool)
    tape = O.noise_tape(1, 25, cfg.ar_vocab())
    best, best_t = 1, float("inf")
    for c in cands:
        torch.set_num_threads(c)
        t0 = time.perf_counter()
        O.ar_generate(sd, cfg, cond, txt, mask, max_frames=24, sampling=O.ArSampling(min_gen_frames=10 ** 9), noise_tv=tape)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
        if dt > 20.0:
            break
    return best, top


def cpu_reference_sample(cfg, sd, n_utts, threads, seed_base=1234):
    """Time the oracle port (reference algorithm, torch CPU eager) on `n_utts` utterances of the
    bench workload, sequentially (the reference has no batch path).  Returns (frames/s, seconds)."""
    from oracle import ar_oracle as O

    torch.set_num_threads(threads)
    cond, txt, _ = _inputs(cfg, 0, n_utts, STEPS_AR, TEXT_LEN)
    samp = O.ArSampling(min_gen_frames=10 ** 9)
    mask = torch.ones(1, TEXT_LEN, dtype=torch.bool)
    t0 = time.perf_counter()
    frames = 0
    for i in range(n_utts):
        tape = O.noise_tape(seed_base + i, STEPS_AR, cfg.ar_vocab())
        toks = O.ar_generate(sd, cfg, cond[i:i + 1], txt[i:i + 1], mask, max_frames=FRAMES, sampling=samp, noise_tv=tape)
        frames += len(toks)
    dt = time.perf_counter() - t0
    return frames / dt, dt


def extras(cfg, dev):
    """Side measurements of the other BASELINE.json configs on one GPU (not the headline `value`):
    batch-1 AR rate (fp32, configs[1]), stream() time-to-first-audio p50 (configs[1]), end-to-end RTF of
    synthesize() at batch 1 and synthesize_batch() at batch 64 (configs[1]/[2]), Mimi decode of 10k frames (configs[4])."""
    from sopro_b200 import SoproTTS
    from sopro_b200.engine import ArEngine, Sampling
    from sopro_b200.tokenizer import IdsTokenizer
    from sopro_b200.weights import synth_mimi_state_dict, synth_state_dict

    out = {}
    sd_full = synth_state_dict(cfg, 1000, 0)
    tts = SoproTTS.from_state_dict(cfg, sd_full, IdsTokenizer(1000), synth_mimi_state_dict(), device=str(dev))
    ref_tokens = torch.randint(0, 2048, (38, 32), generator=torch.Generator().manual_seed(7))
    ref = tts.prepare_reference(ref_tokens_tq=ref_tokens)
    text = " ".join(str(17 * i + 5) for i in range(50))  # 50 ids + BOS/EOS = 52

    def timed(fn, n, warm=2):
        ts = []
        for i in range(n + warm):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize(dev)
            if i >= warm:
                ts.append(time.perf_counter() - t0)
        return float(np.median(ts)), r

    # batch-1 AR rate, fp32 weights, device resident
    eng = tts.model.engine
    cond, txt, noise = _inputs(cfg, 0, 1, STEPS_AR, TEXT_LEN)
    cond, txt, noise = cond.to(dev), txt.to(dev), noise.to(dev)
    ses = eng.session(1, STEPS_AR, TEXT_LEN)
    sp = Sampling(min_gen_frames=2 ** 31 - 1)

    def ar1():
        ses.begin(cond, txt, [TEXT_LEN], noise, sp)
        ses.run()

    t, _ = timed(ar1, 5)
    out["batch1_ar_frames_per_sec_fp32"] = STEPS_AR / t
    out["batch1_us_per_ar_step_fp32"] = t / STEPS_AR * 1e6
    # TTFA p50: stream() with a prepared reference, default chunk_frames=6 (reference streaming.py:141)
    def first_chunk():
        it = tts.stream(text, ref=ref, max_frames=FRAMES, seed=1, min_gen_frames=10 ** 9)
        c = next(it)
        it.close()
        return c

    ts = []
    for i in range(22):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        c = first_chunk()
        torch.cuda.synchronize(dev)
        if i >= 2:
            ts.append(time.perf_counter() - t0)
    out["ttfa_ms_p50"] = float(np.median(ts)) * 1e3
    out["ttfa_first_chunk_samples"] = int(c.numel())
    # RTF: whole synthesize() (tokenize + prefill + AR + NAR + Mimi) / audio seconds
    t, wav = timed(lambda: tts.synthesize(text, ref=ref, max_frames=FRAMES, seed=1, min_gen_frames=10 ** 9), 3, warm=1)
    out["rtf_batch1"] = t / (wav.shape[-1] / 24000.0)
    out["synthesize_batch1_ms"] = t * 1e3
    texts = [" ".join(str(17 * i + 5 + j) for i in range(50)) for j in range(64)]
    # two warm-up calls: the second one is where the (64, 401) NAR shape gets captured into its CUDA graph
    t, wavs = timed(lambda: tts.synthesize_batch(texts, ref=ref, max_frames=FRAMES, seeds=list(range(64)), min_gen_frames=10 ** 9), 3, warm=2)
    out["rtf_batch64"] = t / sum(w.shape[-1] / 24000.0 for w in wavs)
    out["synthesize_batch64_ms"] = t * 1e3
    # Mimi standalone: 25 x 400 = 10k frames
    codes = torch.randint(0, 2048, (25, 32, 400), generator=torch.Generator().manual_seed(5)).to(dev)
    t, _ = timed(lambda: tts.codec.engine.decode(codes), 3, warm=1)
    out["mimi_precision"] = tts.codec.engine.precision  # bf16 operands on tcgen05, fp32 accumulate (default mode)
    out["mimi_frames_per_sec"] = 10000 / t
    out["mimi_ms_per_10k_frames"] = t * 1e3
    out["mimi_alg_gb_per_s"] = 10000 * 7936 / t / 1e9  # codes in + waveform out only
    out["mimi_tflops"] = 10000 * 431.2e6 / t / 1e12     # dense-block FLOPs of one frame: 431.2 MFLOP
    tts.codec.engine.set_precision("fp32")
    t32, _ = timed(lambda: tts.codec.engine.decode(codes), 1, warm=1)
    tts.codec.engine.set_precision("bf16_tc")
    out["mimi_fp32_mode_ms_per_10k_frames"] = t32 * 1e3
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="utterances per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip batch-1 / TTFA / RTF / Mimi side measurements")
    args = ap.parse_args()
    torch.set_grad_enabled(False)
    # Libraries (NCCL's version banner, ...) write to fd 1; the contract is ONE JSON line on stdout.
    # Everything else goes to stderr, the JSON is written to the real stdout at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from sopro_b200.config import SoproTTSConfig
    from sopro_b200.weights import round_through_bf16, synth_state_dict

    cfg = SoproTTSConfig()
    config = {"workload": WORKLOAD, "batch_per_gpu": args.batch, "global_batch": args.batch * max(world, 1),
              "frames": FRAMES, "text_len": TEXT_LEN, "parallelism": f"dp{max(world, 1)}",
              "weights": "synthetic seeded (sopro_b200.weights.synth_state_dict), ar.* rounded through bf16",
              "l2": "L2 flushed (256 MiB write) between timed steps, outside the event pairs"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        sd = round_through_bf16(synth_state_dict(cfg, 64, 0, only_prefix=("ar.", "cb_embed.")), ("ar.", "cb_embed."))
        threads, avail = pick_threads(cfg, sd)
        for _ in range(max(args.warmup, 0)):
            cpu_reference_sample(cfg, sd, 1, threads)  # warm-up: one utterance
        vals, secs = [], 0.0
        for _ in range(max(args.steps, 1)):
            v, dt = cpu_reference_sample(cfg, sd, 1, threads)
            vals.append(v)
            secs += dt
        fps = float(np.sum([STEPS_AR for _ in vals]) / secs)
        line = {"impl": "reference", "metric": "ar_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / len(vals) * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "cores_available": avail, "kind": "port",
                                 "sample": "each step = 1 utterance x 401 AR frames of the bench workload, oracle/ar_oracle.py "
                                           "(torch CPU eager restatement of the reference, bit-equal to it); the reference has no "
                                           "batch path, so batch-64 throughput on CPU is this rate"},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    # ------------------------------------------------------------------ B200 arm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU baseline)")
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from sopro_b200.engine import ArEngine, Sampling

    # weights: built on rank 0, broadcast once over NCCL/NVLink (the only collective of the path)
    sd = None
    if rank == 0:
        sd = round_through_bf16(synth_state_dict(cfg, 64, 0, only_prefix=("ar.", "cb_embed.")), ("ar.", "cb_embed."))
    if world > 1:
        from sopro_b200.dp import broadcast_state_dict
        from sopro_b200.weights import param_specs

        specs = [(k, v[0]) for k, v in param_specs(cfg, 64).items() if k.startswith(("ar.", "cb_embed."))]
        sd = broadcast_state_dict(sd, specs, src=0, device=dev)
    eng = ArEngine(cfg, sd, dev, "bf16")
    B = args.batch
    cond_h, txt_h, noise_h = _inputs(cfg, rank, B, STEPS_AR, TEXT_LEN)
    cond_h, txt_h, noise_h = cond_h.pin_memory(), txt_h.pin_memory(), noise_h.pin_memory()
    cond_d, txt_d, noise_d = cond_h.to(dev), txt_h.to(dev), noise_h.to(dev)
    samp = Sampling(min_gen_frames=2 ** 31 - 1)
    ses = eng.session(B, STEPS_AR, TEXT_LEN)
    lens = [TEXT_LEN] * B
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def one_pass_resident():
        ses.begin(cond_d, txt_d, lens, noise_d, samp)
        ses.run()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        one_pass_resident()
    sync_all()
    clocks = ClockSampler(local_rank)
    clocks.start()
    # ---- device-resident throughput: K passes, each bracketed by its own event pair
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    sync_all()
    for a, m, b in ev:
        flush.fill_(1)
        a.record()
        ses.begin(cond_d, txt_d, lens, noise_d, samp)
        m.record()
        ses.run()
        b.record()
    sync_all()
    t_total_ms = sum(a.elapsed_time(b) for a, _, b in ev)
    t_kernel_ms = sum(m.elapsed_time(b) for _, m, b in ev) / args.steps  # the persistent AR kernel alone
    toks, n_tok, _ = ses.read()
    frames_per_pass = int(n_tok.sum())
    # ---- end to end through the C-ABI host-buffer call: pinned host in, tokens out
    e2e_ms = 0.0
    for i in range(args.steps + 1):
        flush.fill_(1)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        toks_h, n_h = ses.generate_host(cond_h, txt_h, lens, noise_h, samp)
        b.record()
        torch.cuda.synchronize(dev)
        if i > 0:
            e2e_ms += a.elapsed_time(b)
    clocks.stop_flag = True
    clocks.join(timeout=2)
    assert np.array_equal(toks_h, toks), "host-buffer path and resident path disagree"
    t = torch.tensor([t_total_ms, e2e_ms, t_kernel_ms], dtype=torch.float64, device=dev)
    fr = torch.tensor([float(frames_per_pass)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(fr, op=dist.ReduceOp.SUM)
    t_total_ms, e2e_ms, t_kernel_ms = [float(x) for x in t.tolist()]
    frames_all = float(fr.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = frames_all * args.steps / (t_total_ms / 1e3)
    e2e_value = frames_all * args.steps / (e2e_ms / 1e3)
    peak, peak_src = _peaks()
    w_step = eng.step_weight_bytes
    alg_bytes_launch = (w_step + B * S_UTT_BYTES) * STEPS_AR
    achieved = alg_bytes_launch / (t_kernel_ms / 1e3) / 1e9
    traffic = None
    ncu_note = None
    tp = os.path.join(ROOT, "profiles", "ar_kernel_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        traffic = tj.get("dram_bytes_per_launch")
        ncu_note = tj.get("ncu")
    line = {
        "metric": "ar_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": max(world, 1), "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": t_total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 math / bf16 weight storage", "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": "frames/s",
                # whole job: every rank copies its own shard's inputs in and its token ids out
                "h2d_bytes_per_step": int(cond_h.numel() * 4 + txt_h.numel() * 4 + noise_h.numel() * 4) * max(world, 1),
                "d2h_bytes_per_step": int(toks_h.nbytes + n_h.nbytes) * max(world, 1)},
        "gpu_launches": 2 * args.steps * 2 * max(world, 1),  # (kv_build + ar_persistent) per pass, resident + e2e legs, per rank
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "ar_persistent_kernel<bf16>", "ms_per_launch": t_kernel_ms,
                     "alg_bytes_per_launch": alg_bytes_launch, "peak_source": peak_src,
                     "note": "algorithmic bytes = (W_step + B*3280) per AR step x 401 steps (SURVEY.md §8d); W_step is L2-resident "
                             "after the first step, so DRAM traffic is far below this",
                     # last committed ncu --set full capture of this kernel (not measured in this run)
                     "ncu_capture": ncu_note},
        "clocks": clocks.summary(),
        "extra": {"us_per_ar_step": t_kernel_ms / STEPS_AR * 1e3, "frames_per_pass_per_gpu": frames_per_pass},
    }
    if world == 1 and not args.no_extras:
        try:
            line["extra"].update(extras(cfg, dev))
        except Exception as ex:  # the headline number must survive a failure in the side measurements
            line["extra"]["extras_error"] = repr(ex)
    if world > 1:  # the CPU baseline is measured at N=1 only (it does not depend on N)
        line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": None, "kind": "port",
                                "sample": "measured on rank 0 at N=1 only"}
    elif not args.no_cpu_baseline:
        threads, avail = pick_threads(cfg, sd)
        v, dt = cpu_reference_sample(cfg, sd, 16, threads)
        line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": threads, "cores_available": avail, "kind": "port",
                                "sample": f"16 utterances x 401 AR frames of the same workload, sequential, {dt:.1f} s "
                                          "(oracle/ar_oracle.py, torch CPU eager; the reference has no batch path)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

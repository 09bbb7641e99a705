"""First-contact diagnostics on the GPU box: per-block / logits error vs the golden
fixtures, first token divergence, and raw step timings.  Not a test; prints a lot."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import ar_oracle as O
from tests.cases import AR_CASES, ar_case_inputs, _unit
from sopro_b200.engine import ArEngine, Sampling

torch.set_grad_enabled(False)
GOLD = os.path.join(ROOT, "tests", "golden")


def samp_of(s, cfg, **over):
    mg = s.min_gen_frames if s.min_gen_frames is not None else cfg.min_gen_frames
    d = dict(top_p=s.top_p, temperature=s.temperature, anti_loop=s.anti_loop, min_gen_frames=int(min(mg, 2**31 - 1)))
    d.update(over)
    return Sampling(**d)


def check_case(name):
    spec = AR_CASES[name]
    cfg, sd, inp = ar_case_inputs(spec)
    g = np.load(os.path.join(GOLD, f"ar_{name}.npz"))
    eng = ArEngine(cfg, sd, 0, "bf16" if spec["bf16"] else "fp32")
    steps = inp["max_frames"] + 1
    L = inp["txt_seq"].shape[1]
    tape = O.noise_tape(spec["noise_seed"], steps, cfg.ar_vocab())[:, :50].contiguous().unsqueeze(0)
    gold = g["tokens"].tolist()
    # teacher forced
    forced = torch.zeros(1, steps, dtype=torch.int32); forced[0, :len(gold)] = torch.tensor(gold, dtype=torch.int32)
    tr_b = torch.zeros(steps, int(cfg.n_layers_ar), 1, int(cfg.d_model), device="cuda")
    tr_l = torch.zeros(steps, 1, cfg.ar_vocab(), device="cuda")
    ses = eng.session(1, steps, L)
    ses.set_forced(forced); ses.set_trace(tr_b, tr_l)
    ses.begin(inp["cond_ar"], inp["txt_seq"], [L], tape, samp_of(inp["sampling"], cfg))
    ses.run(); toks, n, done = ses.read()
    sampled = ses.sampled().cpu()[0, :len(gold)].tolist()
    bt = g["block_trace"]; got = tr_b[:2, :, 0].cpu().numpy()
    for t in range(2):
        errs = [float(np.abs(got[t, i] - bt[t, i]).max()) for i in range(bt.shape[1])]
        print(f"  [{name}] step {t} block max-abs-err {['%.2e' % e for e in errs]} (ref max {float(np.abs(bt[t]).max()):.2f})")
    lg = tr_l[:, 0].cpu().numpy()
    for row, t in zip(g["logits"], g["logit_steps"].tolist()):
        print(f"  [{name}] logits step {t}: max-abs-err {float(np.abs(lg[t]-row).max()):.2e} (ref max {float(np.abs(row).max()):.2f})")
    mism = [i for i, (a, b) in enumerate(zip(sampled, gold)) if a != b]
    print(f"  [{name}] teacher-forced sampled mismatches: {len(mism)} {mism[:10]}  n={n[0]} done={done[0]}")
    # free running
    ses2 = eng.session(1, steps, L)
    ses2.begin(inp["cond_ar"], inp["txt_seq"], [L], tape, samp_of(inp["sampling"], cfg))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ses2.run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    toks, n, done = ses2.read()
    out = toks[0, :n[0]].tolist()
    first = next((i for i, (a, b) in enumerate(zip(out, gold)) if a != b), None)
    print(f"  [{name}] free-run: len {len(out)} vs {len(gold)}, first divergence {first}, equal={out == gold}, "
          f"{dt*1e3:.2f} ms total, {dt/max(1,n[0])*1e6:.1f} us/step")
    return out == gold


def timing(B, wdtype, steps=401, L=52, team=0):
    spec = AR_CASES["default_bf16" if wdtype == "bf16" else "default_fp32"]
    cfg, sd, _ = ar_case_inputs(spec)
    eng = ArEngine(cfg, sd, 0, wdtype)
    D = int(cfg.d_model)
    cond = _unit(steps * D, 1).view(1, steps, D).expand(B, steps, D).contiguous().cuda()
    cond = cond + 0.01 * torch.arange(B, device="cuda").view(B, 1, 1)
    txt = _unit(L * D, 2).view(1, L, D).expand(B, L, D).contiguous().cuda()
    g = torch.Generator().manual_seed(0)
    noise = torch.empty(B, steps, 50).exponential_(1.0, generator=g).cuda()
    ses = eng.session(B, steps, L)
    if team: ses.set_team(team)
    s = Sampling(min_gen_frames=2**31 - 1)
    res = []
    for it in range(3):
        ses.begin(cond, txt, [L] * B, noise, s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ses.run(); e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1))
    toks, n, done = ses.read()
    ms = min(res)
    print(f"  timing B={B} {wdtype} team={team}: {ms:.2f} ms / {steps} steps = {ms/steps*1e3:.1f} us/step, "
          f"{B*steps/ms*1e3:.0f} frames/s; n={n[:4]} W={eng.step_weight_bytes/1e6:.1f} MB "
          f"algGB/s={(eng.step_weight_bytes + B*3280)*steps/ms/1e6:.0f}")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs")
    names = sys.argv[1:] or list(AR_CASES)
    ok = {}
    for nme in names:
        try:
            ok[nme] = check_case(nme)
        except Exception as ex:  # keep going: first contact wants all the evidence
            import traceback; traceback.print_exc(); ok[nme] = repr(ex)
    print(json.dumps(ok))
    for B, wd, team in [(1, "fp32", 0), (1, "bf16", 0), (8, "fp32", 0), (64, "fp32", 0), (64, "bf16", 0), (64, "bf16", 8), (64, "bf16", 32)]:
        try:
            timing(B, wd, team=team)
        except Exception as ex:
            import traceback; traceback.print_exc()

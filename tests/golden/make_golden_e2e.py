"""Golden fixtures for the stages AROUND the hot path (prefill, NAR), from the reference itself.
Run in the build container only:  python tests/golden/make_golden_e2e.py
The reference modules (unmodified, /root/reference/src) get the synthetic checkpoint; sopro_b200.prefill
runs on the same inputs and must agree; small slices are stored for tests/test_prefill_golden.py."""
import os, sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, "/root/reference/src")
from sopro.config import SoproTTSConfig as RefCfg  # noqa: E402
from sopro.model import SoproTTSModel  # noqa: E402
from sopro_b200 import prefill as P  # noqa: E402
from tests.cases import E2E_CASE, e2e_inputs  # noqa: E402

torch.set_grad_enabled(False)


class _Tok:
    def __init__(self, v): self.vocab_size = v


def main():
    cfg, sd, inp = e2e_inputs()
    ref = SoproTTSModel(RefCfg(), _Tok(E2E_CASE["text_vocab"])).eval()
    ref.load_state_dict(sd, strict=True)
    dev = torch.device("cpu")
    pr = ref.prepare_reference(inp["ref_tokens_tq"], device=dev)
    prep = ref.prepare_conditioning(inp["text_ids"], pr, max_frames=inp["max_frames"], device=dev, style_strength=inp["style_strength"])
    nar = ref.nar_refine(prep["cond_ar"][:, :inp["nar_T"]], inp["rvq1"].unsqueeze(0))
    # ours
    tpos = P.sinusoid_table(int(cfg.max_text_len) + 8, int(cfg.d_model), dev)
    fpos = P.sinusoid_table(int(cfg.pos_emb_max) + 8, int(cfg.d_model), dev)
    mine_pr = P.prepare_reference(sd, cfg, inp["ref_tokens_tq"], dev)
    mine = P.prepare_conditioning(sd, cfg, inp["text_ids"], mine_pr, max_frames=inp["max_frames"], device=dev,
                                  style_strength=inp["style_strength"], text_pos=tpos, frame_pos=fpos)
    mine_nar = P.nar_refine(sd, cfg, mine["cond_ar"][:, :inp["nar_T"]], inp["rvq1"].unsqueeze(0))
    def md(a, b): return float((a - b).abs().max())
    print("sv_ref", md(pr.sv_ref, mine_pr.sv_ref), "ref_seq", md(pr.ref_seq, mine_pr.ref_seq),
          "k0", md(pr.ref_kv_caches[0]["k"], mine_pr.ref_kv_caches[0]["k"]),
          "txt_seq", md(prep["txt_seq"], mine["txt_seq"]), "cond_ar", md(prep["cond_ar"], mine["cond_ar"]),
          "nar equal", bool((nar == mine_nar).all()))
    assert md(pr.sv_ref, mine_pr.sv_ref) < 1e-6 and md(pr.ref_seq, mine_pr.ref_seq) < 1e-5
    assert md(prep["cond_ar"], mine["cond_ar"]) < 1e-5 and md(prep["txt_seq"], mine["txt_seq"]) < 1e-5
    assert bool((nar == mine_nar).all())
    rows = [0, 1, 2, 100, 399, 400]
    np.savez_compressed(os.path.join(HERE, "e2e_prefill.npz"),
                        sv_ref=pr.sv_ref.numpy(), ref_seq_rows=pr.ref_seq[0, :4].numpy(), ref_seq_absmean=np.float32(pr.ref_seq.abs().mean()),
                        k2_rows=pr.ref_kv_caches[2]["k"][0, :, :2].numpy(), txt_seq_rows=prep["txt_seq"][0, :4].numpy(),
                        txt_pool=prep["txt_pool"].numpy(), cond_rows_idx=np.asarray(rows), cond_rows=prep["cond_ar"][0, rows].numpy(),
                        cond_absmean=np.float32(prep["cond_ar"].abs().mean()), nar_tokens=nar[0].numpy().astype(np.int16))
    print("wrote e2e_prefill.npz")


if __name__ == "__main__":
    main()

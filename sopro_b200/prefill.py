"""The stages around the hot path (text encoder, speaker vector, reference encoder + cross-attention, FiLM, NAR refiner)
restated as plain torch ops over the flat state_dict, following the reference's arithmetic op by op (names = the
reference's checkpoint keys).  Since round 2 every one of them runs in a CUDA engine of libsopro_b200.so
(``prefill_cuda.PrefillEngine`` / ``RefPrepEngine``, ``nar.NarEngine``); what the product still takes from this module
is the ``PreparedReference`` dataclass and ``sinusoid_table``.  The functions stay as the CPU restatement the parity
tests check the engines against (itself checked against the unmodified reference by tests/golden/make_golden_e2e.py
and tests/test_host_cpu.py).

Reference map (paths relative to the reference's src/sopro/):
  ssm_block          nn/blocks.py:143-148 (+ DepthwiseConv1d.forward :63-74)
  text_encoder       nn/text.py:29-44
  token2sv           nn/speaker.py:37-61, AttentiveStatsPool nn/blocks.py:174-188
  encode_reference   model.py:133-149
  ref_kv / ref_xattn nn/ref.py:44-108
  film               nn/speaker.py:76-85
  prepare_*          model.py:151-216
  nar_*              nn/nar.py:28-32,89-116, model.py:307-347
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .config import SoproTTSConfig

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass
class PreparedReference:
    """Same fields as the reference's dataclass (model.py:45-50): picklable tensors only."""
    ref_tokens_btq: Tensor
    sv_ref: Tensor
    ref_seq: Tensor
    ref_kv_caches: List[Dict[str, Optional[Tensor]]]


def rms_norm(x: Tensor, w: Tensor, eps: float = 1e-6) -> Tensor:
    x32 = x.float()
    y = x32 * torch.rsqrt(x32.pow(2).mean(dim=-1, keepdim=True) + eps)
    return (y * w.float()).to(x.dtype)


def sinusoid_table(n: int, d: int, device) -> Tensor:
    """nn/embeddings.py:14-22 (non-persistent buffer, rebuilt here)."""
    pe = torch.zeros(n, d)
    pos = torch.arange(0, n, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.to(device)


def dwconv(x_btd: Tensor, w: Tensor, b: Tensor, dilation: int, causal: bool) -> Tensor:
    k = int(w.shape[-1])
    total = (k - 1) * dilation
    left = total if causal else total // 2
    xt = F.pad(x_btd.transpose(1, 2), (left, total - left))
    return F.conv1d(xt, w, b, groups=w.shape[0], dilation=dilation).transpose(1, 2)


def ssm_block(sd: SD, p: str, x: Tensor, dilation: int = 1, causal: bool = False) -> Tensor:
    a, g = F.linear(rms_norm(x, sd[p + "norm.weight"]), sd[p + "glu.pro.weight"], sd[p + "glu.pro.bias"]).chunk(2, dim=-1)
    x = x + dwconv(a * torch.sigmoid(g), sd[p + "dw.dw.weight"], sd[p + "dw.dw.bias"], dilation, causal)
    f = F.linear(rms_norm(x, sd[p + "ff.0.weight"]), sd[p + "ff.1.weight"], sd[p + "ff.1.bias"])
    return x + F.linear(F.gelu(f), sd[p + "ff.3.weight"], sd[p + "ff.3.bias"])


def text_encoder(sd: SD, cfg: SoproTTSConfig, text_ids: Tensor, mask: Tensor, pos_table: Tensor):
    x = sd["text_enc.embed.emb.weight"][text_ids]
    L = x.size(1)
    x = x + pos_table[:L].unsqueeze(0)
    x = x * mask.unsqueeze(-1).float()
    for i in range(int(cfg.n_layers_text)):
        x = ssm_block(sd, f"text_enc.layers.{i}.", x)
    x = rms_norm(x, sd["text_enc.norm.weight"])
    mf = mask.float().unsqueeze(-1)
    return x, (x * mf).sum(dim=1) / (mf.sum(dim=1) + 1e-6)


def token2sv(sd: SD, cfg: SoproTTSConfig, tokens_btq: Tensor, lengths: Optional[Tensor]) -> Tensor:
    B, T, Q = tokens_btq.shape
    dev = tokens_btq.device
    valid = torch.arange(T, device=dev)[None, :] < lengths[:, None] if lengths is not None else torch.ones(B, T, dtype=torch.bool, device=dev)
    idx = torch.arange(Q, device=dev).view(1, 1, Q) * int(cfg.codebook_size) + tokens_btq.long()
    raw = sd["token2sv.emb.weight"][idx] * valid[:, :, None, None].float()
    w = F.softmax(sd["token2sv.cb_weights"], dim=0).view(1, 1, Q, 1)
    x = (raw * w).sum(dim=2) * valid[:, :, None].float()
    h = F.gelu(dwconv(x, sd["token2sv.enc.0.dw.weight"], sd["token2sv.enc.0.dw.bias"], 1, False))
    h = F.gelu(dwconv(h, sd["token2sv.enc.3.dw.weight"], sd["token2sv.enc.3.dw.bias"], 1, False))
    h = h * valid[:, :, None].float()
    logits = F.linear(torch.tanh(F.linear(h, sd["token2sv.pool.attn.0.weight"], sd["token2sv.pool.attn.0.bias"])),
                      sd["token2sv.pool.attn.2.weight"], sd["token2sv.pool.attn.2.bias"]).squeeze(-1)
    if lengths is not None:
        logits = logits.masked_fill(~valid, -1e9)
    a = torch.softmax(logits, dim=1).unsqueeze(-1)
    mu = (h * a).sum(dim=1)
    std = torch.sqrt((a * (h - mu.unsqueeze(1)).pow(2)).sum(dim=1).clamp_min(1e-6))
    e = F.linear(torch.cat([mu, std], dim=-1), sd["token2sv.proj.weight"], sd["token2sv.proj.bias"])
    return F.normalize(e, dim=-1, eps=1e-6)


def encode_reference_seq(sd: SD, cfg: SoproTTSConfig, ref_btq: Tensor) -> Tensor:
    Q, V = int(cfg.num_codebooks), int(cfg.codebook_size)
    w = torch.softmax(sd["ref_cb_weights"].float(), dim=0)
    emb = sd["cb_embed.emb.weight"]
    x = 0.0
    for q in range(Q):
        x = x + w[q] * emb[q * V + ref_btq[:, :, q]]
    for i in range(int(cfg.ref_enc_layers)):
        x = ssm_block(sd, f"ref_enc_blocks.{i}.", x)
    return rms_norm(x, sd["ref_enc_norm.weight"])


def _heads(t: Tensor, h: int) -> Tensor:
    B, T, D = t.shape
    return t.view(B, T, h, D // h).transpose(1, 2)


def ref_kv_caches(sd: SD, cfg: SoproTTSConfig, ref_seq: Tensor) -> List[Dict[str, Optional[Tensor]]]:
    out = []
    H = int(cfg.ref_xattn_heads)
    for i in range(int(cfg.ref_xattn_layers)):
        p = f"ref_xattn.blocks.{i}."
        kv = rms_norm(ref_seq, sd[p + "nkv.weight"])
        out.append({"k": _heads(F.linear(kv, sd[p + "k_proj.weight"]), H), "v": _heads(F.linear(kv, sd[p + "v_proj.weight"]), H),
                    "key_padding_mask": None})
    return out


def ref_xattn(sd: SD, cfg: SoproTTSConfig, x: Tensor, caches) -> Tensor:
    H = int(cfg.ref_xattn_heads)
    for i, c in enumerate(caches):
        p = f"ref_xattn.blocks.{i}."
        q = _heads(F.linear(rms_norm(x, sd[p + "nq.weight"]), sd[p + "q_proj.weight"]), H)
        bias = None
        kpm = c.get("key_padding_mask")
        if kpm is not None:
            kpm = kpm.to(torch.bool)
            bias = torch.zeros((q.size(0), 1, 1, c["k"].size(-2)), device=q.device).masked_fill(kpm[:, None, None, :], float("-inf"))
            bad = kpm.all(dim=1)
            if bad.any():
                bias[bad, :, :, 0] = 0.0
        a = F.scaled_dot_product_attention(q.float(), c["k"].float(), c["v"].float(), attn_mask=bias)
        a = torch.nan_to_num(a, nan=0.0, posinf=0.0, neginf=0.0)
        B, Hh, T, Dh = a.shape
        a = a.transpose(1, 2).contiguous().view(B, T, Hh * Dh)
        rms = lambda t: torch.sqrt(t.float().pow(2).mean(dim=-1, keepdim=True) + 1e-6)  # noqa: E731
        a = (a * (rms(x) / rms(a)).clamp(0.0, 10.0)).to(x.dtype)
        a = F.linear(a, sd[p + "out_proj.weight"])
        x = x + (float(cfg.ref_xattn_gmax) * torch.tanh(sd[p + "gate"])).to(x.dtype) * a
    return x


def film(sd: SD, base: Tensor, sv: Tensor, strength: float) -> Tensor:
    g, b = F.linear(F.gelu(F.linear(sv, sd["spk_film.mlp.0.weight"], sd["spk_film.mlp.0.bias"])),
                    sd["spk_film.mlp.2.weight"], sd["spk_film.mlp.2.bias"]).chunk(2, dim=-1)
    x = F.layer_norm(base, (base.size(-1),), sd["spk_film.norm.weight"], sd["spk_film.norm.bias"])
    return x * (1 + strength * torch.tanh(g.unsqueeze(1))) + strength * torch.tanh(b.unsqueeze(1))


def prepare_reference(sd: SD, cfg: SoproTTSConfig, ref_tokens_tq: Tensor, device) -> PreparedReference:
    ref_btq = ref_tokens_tq.unsqueeze(0).to(device=device, dtype=torch.long)
    lengths = torch.tensor([int(ref_btq.size(1))], device=device, dtype=torch.long)
    sv = token2sv(sd, cfg, ref_btq, lengths)
    seq = encode_reference_seq(sd, cfg, ref_btq)
    return PreparedReference(ref_tokens_btq=ref_btq, sv_ref=sv, ref_seq=seq, ref_kv_caches=ref_kv_caches(sd, cfg, seq))


def prepare_conditioning(sd: SD, cfg: SoproTTSConfig, text_ids_1d: Tensor, ref: PreparedReference, *, max_frames: int,
                         device, style_strength: float, text_pos: Tensor, frame_pos: Tensor) -> Dict[str, Tensor]:
    sv = ref.sv_ref.to(device)
    if sv.dim() == 1:
        sv = sv.unsqueeze(0)
    ids = text_ids_1d.to(device)
    mask = torch.ones_like(ids, dtype=torch.bool).unsqueeze(0)
    txt_seq, txt_pool = text_encoder(sd, cfg, ids.unsqueeze(0), mask, text_pos)
    T = int(max_frames) + 1
    cond = film(sd, txt_pool[:, None, :] + frame_pos[:T].unsqueeze(0), sv, float(style_strength))
    caches = [{k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in c.items()} for c in ref.ref_kv_caches]
    cond = rms_norm(ref_xattn(sd, cfg, cond, caches), sd["cond_norm.weight"])
    return {"txt_seq": txt_seq, "text_mask": mask, "txt_pool": txt_pool, "sv_ref": sv, "cond_ar": cond}


def nar_refine(sd: SD, cfg: SoproTTSConfig, cond_seq: Tensor, rvq1_bt: Tensor, cache: Optional[dict] = None) -> Tensor:
    """`cache` (optional) keeps the per-stage codebook-index tensors on the device between calls, which also makes
    the whole function capturable in a CUDA graph (no host->device copy inside)."""
    B, T, D = cond_seq.shape
    Q, V = int(cfg.num_codebooks), int(cfg.codebook_size)
    out = torch.zeros((B, T, Q), device=cond_seq.device, dtype=torch.long)
    out[:, :, 0] = rvq1_bt
    prev_tok, prev_cb = [rvq1_bt.unsqueeze(-1)], [[0]]
    stages = [(n, idx) for n, idx in cfg.stage_indices().items() if len(idx) > 0]
    emb = sd["cb_embed.emb.weight"]
    dils = cfg.nar_dilations()
    for sid, (name, idxs) in enumerate(stages):
        toks = torch.cat(prev_tok, dim=-1)
        cbs = sum(prev_cb, [])
        ck = ("nar_cbt", sid, str(toks.device))
        cbt = cache.get(ck) if cache is not None else None
        if cbt is None:
            cbt = torch.tensor(cbs, device=toks.device, dtype=torch.long)
            if cache is not None:
                cache[ck] = cbt
        e = emb[cbt.view(1, 1, -1) * V + toks]
        w = F.softmax(sd["nar_prev_cb_weights"].float().index_select(0, cbt), dim=0)
        prev_sum = (e * w.view(1, 1, -1, 1)).sum(dim=2)
        mix = torch.softmax(sd[f"nar.mix.{name}"], dim=0)
        x = mix[0] * cond_seq + mix[1] * prev_sum
        sv = sd["nar.stage_emb.weight"][sid].unsqueeze(0).expand(B, -1)
        g, b = F.linear(F.gelu(F.linear(sv, sd["nar.adapter.mlp.0.weight"], sd["nar.adapter.mlp.0.bias"])),
                        sd["nar.adapter.mlp.2.weight"], sd["nar.adapter.mlp.2.bias"]).chunk(2, dim=-1)
        x = rms_norm(x, sd["nar.adapter.norm.weight"]) * (1 + torch.tanh(g.unsqueeze(1))) + torch.tanh(b.unsqueeze(1))
        for i, d in enumerate(dils):
            x = ssm_block(sd, f"nar.blocks.{i}.", x, dilation=int(d))
        z = F.linear(rms_norm(x, sd["nar.norm.weight"]), sd["nar.pre.weight"], sd["nar.pre.bias"])
        preds = []
        for j in range(len(idxs)):
            hb = sd[f"nar.head_id_emb.{name}.weight"][j].view(1, 1, -1)
            preds.append(F.linear(z + hb, sd[f"nar.heads.{name}.{j}.weight"], sd[f"nar.heads.{name}.{j}.bias"]).argmax(dim=-1))
        preds = torch.stack(preds, dim=-1)
        for j, cb in enumerate(idxs):
            out[:, :, cb] = preds[:, :, j]
        prev_tok.append(preds)
        prev_cb.append(list(idxs))
    return out

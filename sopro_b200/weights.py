"""Parameter inventory, synthetic checkpoints and weight loading.

The tensor names and shapes are those of the reference's ``state_dict`` (what
``model.safetensors`` holds; reference: src/sopro/model.py:53-117 and the
sub-modules it instantiates), so a real checkpoint loads unchanged.

There is no network here, hence no real checkpoint: ``synth_state_dict`` builds
a seeded synthetic one.  It deliberately does NOT use torch's RNG: values come
from a counter-based integer hash (splitmix64) turned into uniforms with exact
float arithmetic, so the same bytes are produced on any host (this container,
the GPU box) and the golden fixtures under tests/golden/ stay valid.
"""
from __future__ import annotations

import json
import math
import struct
import zlib
from collections import OrderedDict
from typing import Dict, Iterable, Optional, Tuple

import numpy as np
import torch

from .config import SoproTTSConfig

# kind -> how the synthetic generator fills it
#   "lin"   uniform(+-1/sqrt(fan_in))      (torch Linear / Conv default bound)
#   "bias"  uniform(+-1/sqrt(fan_in))
#   "emb"   uniform(+-sqrt(3))             (unit variance, like nn.Embedding's N(0,1))
#   "norm"  uniform(0.8, 1.2)              (live, not all-ones)
#   "small" uniform(+-0.02*sqrt(3))        (layers the reference zero-inits; made live)
#   "gate"  constant 0.5                   (reference inits gates to 0 -> dead branch)
#   "mix"   uniform(+-0.5)
#   "lins"  linspace(1.0, 0.1, n)          (reference: model.py:113-117, speaker.py:22-23)


def _ssm_block(specs, p: str, d: int, k: int):
    specs[p + "norm.weight"] = ((d,), "norm", d)
    specs[p + "glu.pro.weight"] = ((2 * d, d), "lin", d)
    specs[p + "glu.pro.bias"] = ((2 * d,), "bias", d)
    specs[p + "dw.dw.weight"] = ((d, 1, k), "lin", k)
    specs[p + "dw.dw.bias"] = ((d,), "bias", k)
    specs[p + "ff.0.weight"] = ((d,), "norm", d)
    specs[p + "ff.1.weight"] = ((4 * d, d), "lin", d)
    specs[p + "ff.1.bias"] = ((4 * d,), "bias", d)
    specs[p + "ff.3.weight"] = ((d, 4 * d), "lin", 4 * d)
    specs[p + "ff.3.bias"] = ((d,), "bias", 4 * d)


def _xattn(specs, p: str, d: int):
    specs[p + "gate"] = ((), "gate", 1)
    specs[p + "nq.weight"] = ((d,), "norm", d)
    specs[p + "nkv.weight"] = ((d,), "norm", d)
    for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
        specs[p + n + ".weight"] = ((d, d), "lin", d)


def param_specs(cfg: SoproTTSConfig, text_vocab: int) -> "OrderedDict[str, tuple]":
    """name -> (shape, kind, fan_in) for every tensor of the reference state_dict."""
    D = int(cfg.d_model)
    Q, V = int(cfg.num_codebooks), int(cfg.codebook_size)
    sv = int(cfg.sv_student_dim)
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["nar_prev_cb_weights"] = ((Q,), "mix", 1)
    # text encoder (reference: nn/text.py:16-27)
    s["text_enc.embed.emb.weight"] = ((int(text_vocab), D), "emb", 1)
    for i in range(int(cfg.n_layers_text)):
        _ssm_block(s, f"text_enc.layers.{i}.", D, 7)
    s["text_enc.norm.weight"] = ((D,), "norm", D)
    # codebook embedding (reference: nn/embeddings.py:37-49)
    s["cb_embed.emb.weight"] = ((Q * V + 1, D), "emb", 1)
    # speaker vector student (reference: nn/speaker.py:12-30)
    s["token2sv.cb_weights"] = ((Q,), "lins", 1)
    s["token2sv.emb.weight"] = ((Q * V, 192), "emb", 1)
    for i in (0, 3):
        s[f"token2sv.enc.{i}.dw.weight"] = ((192, 1, 7), "lin", 7)
        s[f"token2sv.enc.{i}.dw.bias"] = ((192,), "bias", 7)
    s["token2sv.pool.attn.0.weight"] = ((192, 192), "lin", 192)
    s["token2sv.pool.attn.0.bias"] = ((192,), "bias", 192)
    s["token2sv.pool.attn.2.weight"] = ((1, 192), "lin", 192)
    s["token2sv.pool.attn.2.bias"] = ((1,), "bias", 192)
    s["token2sv.proj.weight"] = ((sv, 384), "lin", 384)
    s["token2sv.proj.bias"] = ((sv,), "bias", 384)
    # FiLM (reference: nn/speaker.py:64-74); last layer zero-init there -> "small"
    s["spk_film.mlp.0.weight"] = ((D, sv), "lin", sv)
    s["spk_film.mlp.0.bias"] = ((D,), "bias", sv)
    s["spk_film.mlp.2.weight"] = ((2 * D, D), "small", D)
    s["spk_film.mlp.2.bias"] = ((2 * D,), "small", D)
    s["spk_film.norm.weight"] = ((D,), "norm", D)
    s["spk_film.norm.bias"] = ((D,), "small", D)
    # AR generator (reference: nn/generator.py:12-42)
    attn = set(cfg.ar_attn_layers())
    for i in range(int(cfg.n_layers_ar)):
        _ssm_block(s, f"ar.blocks.{i}.", D, int(cfg.ar_kernel))
    for i in range(int(cfg.n_layers_ar)):
        if i in attn:
            _xattn(s, f"ar.x_attns.{i}.", D)
    s["ar.norm.weight"] = ((D,), "norm", D)
    s["ar.head.weight"] = ((cfg.ar_vocab(), D), "lin", D)
    s["ar.head.bias"] = ((cfg.ar_vocab(),), "bias", D)
    # NAR refiner (reference: nn/nar.py:35-86)
    for i in range(int(cfg.n_layers_nar)):
        _ssm_block(s, f"nar.blocks.{i}.", D, int(cfg.nar_kernel_size))
    Hn = int(cfg.nar_head_dim)
    stages = [(n, idx) for n, idx in cfg.stage_indices().items() if len(idx) > 0]
    s["nar.norm.weight"] = ((D,), "norm", D)
    s["nar.pre.weight"] = ((Hn, D), "lin", D)
    s["nar.pre.bias"] = ((Hn,), "bias", D)
    s["nar.stage_emb.weight"] = ((len(stages), D), "emb", 1)
    s["nar.adapter.norm.weight"] = ((D,), "norm", D)
    s["nar.adapter.mlp.0.weight"] = ((256, D), "lin", D)
    s["nar.adapter.mlp.0.bias"] = ((256,), "bias", D)
    s["nar.adapter.mlp.2.weight"] = ((2 * D, 256), "small", 256)
    s["nar.adapter.mlp.2.bias"] = ((2 * D,), "small", 256)
    for n, idx in stages:
        for j in range(len(idx)):
            s[f"nar.heads.{n}.{j}.weight"] = ((V, Hn), "lin", Hn)
            s[f"nar.heads.{n}.{j}.bias"] = ((V,), "bias", Hn)
    for n, idx in stages:
        s[f"nar.head_id_emb.{n}.weight"] = ((len(idx), Hn), "small", 1)
    for n, idx in stages:
        s[f"nar.mix.{n}"] = ((2,), "mix", 1)
    s["cond_norm.weight"] = ((D,), "norm", D)
    # reference encoder + cross-attention (reference: model.py:100-117, nn/ref.py)
    for i in range(int(cfg.ref_enc_layers)):
        _ssm_block(s, f"ref_enc_blocks.{i}.", D, 7)
    s["ref_enc_norm.weight"] = ((D,), "norm", D)
    for i in range(int(cfg.ref_xattn_layers)):
        _xattn(s, f"ref_xattn.blocks.{i}.", D)
    s["ref_cb_weights"] = ((Q,), "lins", 1)
    return s


def ar_step_param_names(cfg: SoproTTSConfig) -> Tuple[str, ...]:
    """Names of the tensors the AR step reads every frame (SURVEY.md §8d W_step)."""
    names = []
    attn = set(cfg.ar_attn_layers())
    for i in range(int(cfg.n_layers_ar)):
        p = f"ar.blocks.{i}."
        names += [p + n for n in ("norm.weight", "glu.pro.weight", "glu.pro.bias",
                                  "dw.dw.weight", "dw.dw.bias", "ff.0.weight",
                                  "ff.1.weight", "ff.1.bias", "ff.3.weight", "ff.3.bias")]
        if i in attn:
            q = f"ar.x_attns.{i}."
            names += [q + n for n in ("nq.weight", "q_proj.weight", "out_proj.weight", "gate")]
    names += ["ar.norm.weight", "ar.head.weight", "ar.head.bias"]
    return tuple(names)


# ---------------------------------------------------------------------------
# platform-independent synthetic values
# ---------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        z = z ^ (z >> np.uint64(31))
    return z


def hash_uniform(n: int, key: int) -> np.ndarray:
    """n float32 uniforms in [-1, 1) from integer hashing only (exact in fp32)."""
    with np.errstate(over="ignore"):
        base = _splitmix64(np.array([key], dtype=np.uint64))[0]
        ctr = np.arange(n, dtype=np.uint64) + base
    h = _splitmix64(ctr)
    m = (h >> np.uint64(40)).astype(np.int64)  # 24 random bits
    # (m - 2^23) * 2^-23 is exact in fp32
    return ((m - (1 << 23)).astype(np.float32)) * np.float32(2.0 ** -23)


def _name_key(name: str, seed: int) -> int:
    return (zlib.crc32(name.encode("utf-8")) << 20) ^ (int(seed) & 0xFFFFF)


def synth_tensor(name: str, shape, kind: str, fan_in: int, seed: int) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    if kind == "gate":
        a = np.full((n,), 0.5, dtype=np.float32)
    elif kind == "lins":
        a = torch.linspace(1.0, 0.1, n).numpy().astype(np.float32)
    else:
        u = hash_uniform(n, _name_key(name, seed))
        if kind in ("lin", "bias"):
            a = u * np.float32(1.0 / np.sqrt(float(fan_in)))
        elif kind == "emb":
            a = u * np.float32(np.sqrt(3.0))
        elif kind == "norm":
            a = np.float32(1.0) + u * np.float32(0.2)
        elif kind == "small":
            a = u * np.float32(0.02 * np.sqrt(3.0))
        elif kind == "mix":
            a = u * np.float32(0.5)
        else:
            raise ValueError(kind)
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.reshape(tuple(shape))


def synth_state_dict(
    cfg: SoproTTSConfig,
    text_vocab: int = 128257,
    seed: int = 0,
    *,
    only_prefix: Optional[Iterable[str]] = None,
    head_gain: float = 1.0,
) -> Dict[str, torch.Tensor]:
    """Seeded synthetic checkpoint with every branch live (see module docstring).

    ``only_prefix`` restricts generation to names starting with one of the given
    prefixes (e.g. ("ar.", "cb_embed.") for AR-only tests).  ``head_gain``
    scales ``ar.head.weight`` so the logits are not pessimistically flat.
    """
    out: Dict[str, torch.Tensor] = {}
    pref = tuple(only_prefix) if only_prefix is not None else None
    for name, (shape, kind, fan_in) in param_specs(cfg, text_vocab).items():
        if pref is not None and not name.startswith(pref):
            continue
        t = synth_tensor(name, shape, kind, fan_in, seed)
        if name == "ar.head.weight" and head_gain != 1.0:
            t = t * float(head_gain)
        out[name] = t
    return out


def round_through_bf16(sd: Dict[str, torch.Tensor], prefixes=("ar.",)) -> Dict[str, torch.Tensor]:
    """fp32 state_dict whose ``prefixes`` tensors hold bf16-representable values.

    This is the weight set both the bf16 engine and its oracle use
    (the oracle must see the same rounded values)."""
    out = {}
    for k, v in sd.items():
        if k.startswith(tuple(prefixes)) and v.is_floating_point():
            out[k] = v.to(torch.bfloat16).to(torch.float32)
        else:
            out[k] = v
    return out


# ---------------------------------------------------------------------------
# safetensors (reference: src/sopro/hub.py:30-52)
# ---------------------------------------------------------------------------
def read_safetensors_cfg(path: str) -> SoproTTSConfig:
    with open(path, "rb") as f:
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen).decode("utf-8"))
    meta = header.get("__metadata__", {}) or {}
    if "cfg" not in meta:
        raise RuntimeError(f"No 'cfg' metadata found in {path}.")
    return SoproTTSConfig.from_dict(json.loads(meta["cfg"]))


def load_safetensors(path: str) -> Dict[str, torch.Tensor]:
    from safetensors.torch import load_file

    return load_file(path)


def save_safetensors(sd: Dict[str, torch.Tensor], cfg: SoproTTSConfig, path: str) -> None:
    from safetensors.torch import save_file

    save_file({k: v.contiguous() for k, v in sd.items()}, path, metadata={"cfg": cfg.to_json()})


# ---------------------------------------------------------------------------
# synthetic Mimi decode-path weights (no kyutai/mimi checkpoint offline)
# ---------------------------------------------------------------------------
def synth_mimi_state_dict(seed: int = 5, layer_scale: float = 0.3) -> Dict[str, torch.Tensor]:
    """Seeded random weights for the decode path of ``MimiModel(MimiConfig(num_quantizers=32))``, platform
    independent (integer hashing, like sopro_b200.weights).  LayerScale is raised from its 0.01 init so the
    attention / MLP branches are visible in the output."""
    def U(name, shape, bound):
        import zlib
        n = int(np.prod(shape))
        return torch.from_numpy(hash_uniform(n, (zlib.crc32(name.encode()) << 20) ^ seed) * np.float32(bound)).view(shape)

    sd: Dict[str, torch.Tensor] = {}
    for grp, n in (("semantic", 1), ("acoustic", 31)):
        pre = f"quantizer.{grp}_residual_vector_quantizer."
        for i in range(n):
            sd[pre + f"layers.{i}.codebook.embed_sum"] = U(pre + f"{i}.e", (2048, 256), 1.0)
            sd[pre + f"layers.{i}.codebook.cluster_usage"] = U(pre + f"{i}.u", (2048,), 0.4) + 1.0
            sd[pre + f"layers.{i}.codebook.initialized"] = torch.ones(1)
        sd[pre + "output_proj.weight"] = U(pre + "o", (512, 256, 1), 1 / 16.0)
        sd[pre + "input_proj.weight"] = U(pre + "i", (256, 512, 1), 1 / 22.6)
    sd["upsample.conv.weight"] = U("up", (512, 1, 4), 0.7)
    for l in range(8):
        p = f"decoder_transformer.layers.{l}."
        for n in ("q", "k", "v", "o"):
            sd[p + f"self_attn.{n}_proj.weight"] = U(p + n, (512, 512), 1 / 22.6)
        sd[p + "mlp.fc1.weight"] = U(p + "f1", (2048, 512), 1 / 22.6)
        sd[p + "mlp.fc2.weight"] = U(p + "f2", (512, 2048), 1 / 45.0)
        for n in ("input_layernorm", "post_attention_layernorm"):
            sd[p + n + ".weight"] = 1.0 + U(p + n + "w", (512,), 0.2)
            sd[p + n + ".bias"] = U(p + n + "b", (512,), 0.1)
        sd[p + "self_attn_layer_scale.scale"] = layer_scale + U(p + "ls1", (512,), 0.1)
        sd[p + "mlp_layer_scale.scale"] = layer_scale + U(p + "ls2", (512,), 0.1)
    chans = [(512, 1024, 7)]
    sd["decoder.layers.0.conv.weight"] = U("d0w", (1024, 512, 7), 1 / math.sqrt(512 * 7))
    sd["decoder.layers.0.conv.bias"] = U("d0b", (1024,), 0.02)
    li, c = 1, 1024
    for r in (8, 6, 5, 4):
        sd[f"decoder.layers.{li + 1}.conv.weight"] = U(f"t{li}w", (c, c // 2, 2 * r), 1 / math.sqrt(c * 2))
        sd[f"decoder.layers.{li + 1}.conv.bias"] = U(f"t{li}b", (c // 2,), 0.02)
        p = f"decoder.layers.{li + 2}.block."
        sd[p + "1.conv.weight"] = U(p + "1w", (c // 4, c // 2, 3), 1 / math.sqrt(c // 2 * 3))
        sd[p + "1.conv.bias"] = U(p + "1b", (c // 4,), 0.02)
        sd[p + "3.conv.weight"] = U(p + "3w", (c // 2, c // 4, 1), 1 / math.sqrt(c // 4))
        sd[p + "3.conv.bias"] = U(p + "3b", (c // 2,), 0.02)
        li += 3
        c //= 2
    sd[f"decoder.layers.{li + 1}.conv.weight"] = U("lw", (1, 64, 3), 1 / math.sqrt(64 * 3))
    sd[f"decoder.layers.{li + 1}.conv.bias"] = U("lb", (1,), 0.02)
    return sd


def synth_mimi_encoder_state_dict(seed: int = 5, layer_scale: float = 0.3) -> Dict[str, torch.Tensor]:
    """Seeded random weights for the ENCODE path of ``MimiModel`` (SEANet encoder, encoder transformer, the
    replicate-padded downsampling conv); the quantizer's codebooks and input projections come from
    ``synth_mimi_state_dict``.  Same hashing, same platform independence."""
    def U(name, shape, bound):
        import zlib
        n = int(np.prod(shape))
        return torch.from_numpy(hash_uniform(n, (zlib.crc32(name.encode()) << 20) ^ seed) * np.float32(bound)).view(shape)

    sd: Dict[str, torch.Tensor] = {}
    sd["encoder.layers.0.conv.weight"] = U("e0w", (64, 1, 7), 1.5)
    sd["encoder.layers.0.conv.bias"] = U("e0b", (64,), 0.05)
    li, c = 1, 64
    for r in (4, 5, 6, 8):  # reversed(upsampling_ratios), modeling_mimi.py:465
        p = f"encoder.layers.{li}.block."
        sd[p + "1.conv.weight"] = U(p + "1w", (c // 2, c, 3), 1.7 / math.sqrt(c * 3))
        sd[p + "1.conv.bias"] = U(p + "1b", (c // 2,), 0.02)
        sd[p + "3.conv.weight"] = U(p + "3w", (c, c // 2, 1), 1.7 / math.sqrt(c // 2))
        sd[p + "3.conv.bias"] = U(p + "3b", (c,), 0.02)
        sd[f"encoder.layers.{li + 2}.conv.weight"] = U(f"ed{li}w", (2 * c, c, 2 * r), 2.5 / math.sqrt(c * 2 * r))
        sd[f"encoder.layers.{li + 2}.conv.bias"] = U(f"ed{li}b", (2 * c,), 0.02)
        li += 3
        c *= 2
    sd[f"encoder.layers.{li + 1}.conv.weight"] = U("elw", (512, c, 3), 2.5 / math.sqrt(c * 3))
    sd[f"encoder.layers.{li + 1}.conv.bias"] = U("elb", (512,), 0.02)
    for l in range(8):
        p = f"encoder_transformer.layers.{l}."
        for n in ("q", "k", "v", "o"):
            sd[p + f"self_attn.{n}_proj.weight"] = U(p + n, (512, 512), 1 / 22.6)
        sd[p + "mlp.fc1.weight"] = U(p + "f1", (2048, 512), 1 / 22.6)
        sd[p + "mlp.fc2.weight"] = U(p + "f2", (512, 2048), 1 / 45.0)
        for n in ("input_layernorm", "post_attention_layernorm"):
            sd[p + n + ".weight"] = 1.0 + U(p + n + "w", (512,), 0.2)
            sd[p + n + ".bias"] = U(p + n + "b", (512,), 0.1)
        sd[p + "self_attn_layer_scale.scale"] = layer_scale + U(p + "ls1", (512,), 0.1)
        sd[p + "mlp_layer_scale.scale"] = layer_scale + U(p + "ls2", (512,), 0.1)
    sd["downsample.conv.weight"] = U("down", (512, 512, 4), 2.0 / math.sqrt(512 * 4))
    return sd

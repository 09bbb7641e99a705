"""Design check (CPU) for the planned fused SEANet stage kernel.

The fused kernel keeps everything in the *input-row* layout: row i of the ConvTranspose input owns the r output rows
i*r .. i*r + r-1 as r "phases" of `cout` channels.  In that layout the ResnetBlock's k=3 conv becomes ONE GEMM with a
block-banded weight over [previous row's last two phases | this row's r phases], and its 1x1 conv a block-diagonal
GEMM that accumulates onto z (the skip).  This test builds those packed weights from the state_dict tensors and checks
the three GEMMs against the oracle's layer-by-layer computation of the same stage, so that the index algebra is
settled before any CUDA is written.  It also pins the packing the current engine uses for the ConvTranspose
(mimi_engine.cu: tap 0 <-> x[t-1] <-> w[.., phase + r], tap 1 <-> x[t] <-> w[.., phase]).
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import mimi_oracle as M

torch.set_grad_enabled(False)


def pack_convt(wt, r):
    """ConvTranspose1d weight [cin, cout, 2r] -> [r*cout, 2*cin]: column block 0 multiplies x[i-1], block 1 x[i]."""
    cin, cout, _ = wt.shape
    w = torch.zeros(r * cout, 2 * cin)
    for p in range(r):
        w[p * cout:(p + 1) * cout, :cin] = wt[:, :, p + r].t()
        w[p * cout:(p + 1) * cout, cin:] = wt[:, :, p].t()
    return w


def pack_res1_banded(w1, r):
    """Conv1d weight [hid, cout, 3] -> [r*hid, (r+2)*cout] over [prev row phases r-2, r-1 | this row phases 0..r-1]."""
    hid, cout, k = w1.shape
    assert k == 3
    w = torch.zeros(r * hid, (r + 2) * cout)
    for pp in range(r):           # output phase
        for j in range(3):        # tap j reads output row o + j - 2  ->  block index bq = pp + j
            bq = pp + j
            w[pp * hid:(pp + 1) * hid, bq * cout:(bq + 1) * cout] = w1[:, :, j]
    return w


def pack_res2_blockdiag(w2, r):
    """Conv1d 1x1 weight [cout, hid, 1] -> block diagonal [r*cout, r*hid]."""
    cout, hid, _ = w2.shape
    w = torch.zeros(r * cout, r * hid)
    for p in range(r):
        w[p * cout:(p + 1) * cout, p * hid:(p + 1) * hid] = w2[:, :, 0]
    return w


@pytest.mark.parametrize("stage", [2, 3])  # the 5x and 4x stages (0-based: ratios 8, 6, 5, 4)
def test_input_row_layout_reproduces_a_seanet_stage(stage):
    sd = M.synth_mimi_state_dict()
    r = M.UPSAMPLING_RATIOS[stage]
    li = 1 + 3 * stage
    wt, bt = sd[f"decoder.layers.{li + 1}.conv.weight"], sd[f"decoder.layers.{li + 1}.conv.bias"]
    p = f"decoder.layers.{li + 2}.block."
    w1, b1, w2, b2 = sd[p + "1.conv.weight"], sd[p + "1.conv.bias"], sd[p + "3.conv.weight"], sd[p + "3.conv.bias"]
    cin, cout = wt.shape[0], wt.shape[1]
    hid = w1.shape[0]
    T = 11
    x = torch.randn(1, T, cin, generator=torch.Generator().manual_seed(stage))  # the stage's input BEFORE its ELU
    # ---- oracle, layer by layer (channel-last)
    z_ref = M.conv_transpose_causal(F.elu(x), wt, bt, r)
    h_ref = M.conv1d_causal(F.elu(z_ref), w1, b1)
    out_ref = z_ref + M.conv1d_causal(F.elu(h_ref), w2, b2)
    # ---- input-row layout, three GEMMs
    a = F.elu(x)[0]                                              # [T, cin]
    a_prev = torch.cat([torch.zeros(1, cin), a[:-1]], dim=0)     # x[i-1], zero before the start (causal)
    z = torch.cat([a_prev, a], dim=1) @ pack_convt(wt, r).t() + bt.repeat(r)          # [T, r*cout]
    e = F.elu(z)
    e_prev_tail = torch.cat([torch.zeros(1, 2 * cout), e[:-1, (r - 2) * cout:]], dim=0)  # previous row's last 2 phases
    h = torch.cat([e_prev_tail, e], dim=1) @ pack_res1_banded(w1, r).t() + b1.repeat(r)  # [T, r*hid]
    out = z + F.elu(h) @ pack_res2_blockdiag(w2, r).t() + b2.repeat(r)                 # accumulates onto z: the skip
    got = out.view(T * r, cout)
    scale = float(out_ref.abs().max())
    assert float((z.view(T * r, cout) - z_ref[0]).abs().max()) <= 1e-5 * max(1.0, float(z_ref.abs().max()))
    assert float((h.view(T * r, hid) - h_ref[0]).abs().max()) <= 1e-5 * max(1.0, float(h_ref.abs().max()))
    assert float((got - out_ref[0]).abs().max()) <= 1e-5 * max(1.0, scale)

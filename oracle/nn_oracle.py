"""CPU restatement of the PointNet-attention forward -- TEST INFRASTRUCTURE ONLY.

Plain fp32 torch tensor algebra on the CPU (matmul / mean / var / softmax written out), driven by a
`state_dict` with the reference's parameter names. Each function cites the reference lines it follows
(paths relative to the reference repo):

  tnet()            pointNet/model/pointnetAtt.py:28-47    TransformationNet.forward
  base_pointnet()   pointNet/model/pointnetAtt.py:80-112   BasePointNet.forward
  seg_attention()   pointNet/model/pointnetAtt.py:176-209  SegmentationWithAttention.forward
  train_step_loss() pointNet/self-attention/train_pointnet-attention.py:445-465 (CE + 0.001 * reg)

The arithmetic dependency of the reference for this stage is PyTorch itself (nn.Conv1d(k=1),
BatchNorm1d, MaxPool1d, Linear, MultiheadAttention); their published semantics are restated here and
PINNED against the unmodified reference modules executed in-container with a shared state_dict
(tests/test_oracle_pinned.py::test_nn_oracle_matches_reference_live, tests/golden/nn_reference.npz made
by oracle/make_golden_nn.py). Gradients for the backward parity come from autograd through these
functions. Dropout: the oracle takes explicit keep-masks (None = dropout off), because the reference's
random masks cannot be reproduced bit-for-bit; parity is defined with dropout off.
"""
import math

import torch

EPS = 1e-5  # nn.BatchNorm1d default


def _bn(y, sd, pre, training, stats=None):
    """BatchNorm1d over all rows of y [M, C] (reference: bn_k(conv_k(x)) on [B, C, N] == per-channel over B*N).
    training: batch statistics (biased variance) as nn.BatchNorm1d.train(); when `stats` (a dict) is given
    the running statistics held in `sd` are updated IN PLACE with momentum 0.1 (unbiased variance) and
    num_batches_tracked is incremented, as the module does, and stats[pre] counts the calls.
    ATen's batch_norm is the reference's own arithmetic here (NN arithmetic dependency = PyTorch)."""
    w, b = sd[pre + ".weight"], sd[pre + ".bias"]
    if not training:
        return torch.nn.functional.batch_norm(y, sd[pre + ".running_mean"], sd[pre + ".running_var"], w, b, False, 0.1, EPS)
    if stats is None:
        return torch.nn.functional.batch_norm(y, None, None, w, b, True, 0.1, EPS)
    stats[pre] = stats.get(pre, 0) + 1
    sd[pre + ".num_batches_tracked"] += 1
    return torch.nn.functional.batch_norm(y, sd[pre + ".running_mean"], sd[pre + ".running_var"], w, b, True, 0.1, EPS)


def _conv(x, sd, pre):
    """nn.Conv1d(cin, cout, 1) on [B, cin, N] == x[M, cin] @ W[cout, cin]^T (+ bias)."""
    w = sd[pre + ".weight"]
    w = w.reshape(w.shape[0], -1)
    y = x @ w.t()
    b = sd.get(pre + ".bias")
    return y if b is None else y + b


def tnet(sd, pre, x, training=False, stats=None):
    """x [B, N, d] -> [B, d, d]   (pointnetAtt.py:28-47)"""
    B, N, d = x.shape
    h = x.reshape(B * N, d)
    h = torch.relu(_bn(_conv(h, sd, pre + "conv_1"), sd, pre + "bn_1", training, stats))     # :31
    h = torch.relu(_bn(_conv(h, sd, pre + "conv_2"), sd, pre + "bn_2", training, stats))     # :32
    h = torch.relu(_bn(_conv(h, sd, pre + "conv_3"), sd, pre + "bn_3", training, stats))     # :33
    g = h.reshape(B, N, 256).max(dim=1).values                                               # :35-36
    g = torch.relu(_bn(g @ sd[pre + "fc_1.weight"].t(), sd, pre + "bn_4", training, stats))  # :38
    g = torch.relu(_bn(g @ sd[pre + "fc_2.weight"].t(), sd, pre + "bn_5", training, stats))  # :39
    g = g @ sd[pre + "fc_3.weight"].t() + sd[pre + "fc_3.bias"]                              # :40
    return g.reshape(B, d, d) + torch.eye(d, dtype=x.dtype)                                  # :42-46


def base_pointnet(sd, x, training=False, stats=None, point_dimension=3):
    """x [B, N, 9] -> (out [B, N, 320] = [global x N | local 64], feature_transform [B, 64, 64])  (:80-112)"""
    B, N, _ = x.shape
    xt = x[:, :, :point_dimension]                                                            # :83
    T = tnet(sd, "input_transform.", xt, training, stats)                                    # :84
    xt = torch.bmm(xt, T)                                                                     # :85
    h = torch.cat([xt, x], dim=2).reshape(B * N, -1)                                          # :86
    h = torch.relu(_bn(_conv(h, sd, "conv_1"), sd, "bn_1", training, stats))                  # :90
    h = torch.relu(_bn(_conv(h, sd, "conv_2"), sd, "bn_2", training, stats))                  # :91
    F64 = tnet(sd, "feature_transform.", h.reshape(B, N, 64), training, stats)                # :94
    local = torch.bmm(h.reshape(B, N, 64), F64)                                               # :96-97
    h = local.reshape(B * N, 64)
    h = torch.relu(_bn(_conv(h, sd, "conv_3"), sd, "bn_3", training, stats))                  # :100
    h = torch.relu(_bn(_conv(h, sd, "conv_4"), sd, "bn_4", training, stats))                  # :101
    h = torch.relu(_bn(_conv(h, sd, "conv_5"), sd, "bn_5", training, stats))                  # :102
    h = torch.relu(_bn(_conv(h, sd, "conv_6"), sd, "bn_6", training, stats))                  # :103
    g = h.reshape(B, N, -1).max(dim=1).values                                                 # :104-106
    out = torch.cat([g[:, None, :].expand(B, N, g.shape[1]), local], dim=2)                   # :109-110
    return out, F64


def mha(sd, pre, x, num_heads, key_padding_mask=None, attn_keep=None):
    """nn.MultiheadAttention(E, H) self-attention, seq-first x [L, B, E] -> ([L, B, E], weights [B, L, L]).
    attn_keep: optional [B*H, L, L] keep-mask already scaled by 1/(1-p) (dropout on the softmax output)."""
    L, B, E = x.shape
    hd = E // num_heads
    qkv = x @ sd[pre + "in_proj_weight"].t() + sd[pre + "in_proj_bias"]
    q, k, v = qkv.split(E, dim=-1)
    q = q.reshape(L, B * num_heads, hd).transpose(0, 1) * (1.0 / math.sqrt(hd))
    k = k.reshape(L, B * num_heads, hd).transpose(0, 1)
    v = v.reshape(L, B * num_heads, hd).transpose(0, 1)
    s = torch.bmm(q, k.transpose(1, 2))                          # [B*H, L, L]
    if key_padding_mask is not None:
        m = key_padding_mask[:, None, None, :].expand(B, num_heads, 1, L).reshape(B * num_heads, 1, L)
        s = s.masked_fill(m, float("-inf"))
    p = torch.softmax(s, dim=-1)
    pd = p if attn_keep is None else p * attn_keep
    o = torch.bmm(pd, v).transpose(0, 1).reshape(L, B, E)
    o = o @ sd[pre + "out_proj.weight"].t() + sd[pre + "out_proj.bias"]
    return o, p.reshape(B, num_heads, L, L).mean(1)


def seg_attention(sd, gl_feats, lo_feats, centroids, np_cluster, attn_mask=None, training=False, stats=None,
                  num_heads=8, attn_keep=None, keep_2=None, keep_3=None):
    """gl_feats [W, B, E], lo_feats [B, sumN, 64], centroids [B, W, 2] -> logits [B, C, sumN]  (:176-209)"""
    pe = torch.nn.functional.leaky_relu(centroids @ sd["fc1.weight"].t() + sd["fc1.bias"])   # :183
    pe = pe @ sd["fc2.weight"].t() + sd["fc2.bias"]
    g = gl_feats + pe.transpose(0, 1)                                                          # :184-185
    g, _ = mha(sd, "attention.", g, num_heads, attn_mask, attn_keep)                          # :187-190
    B = lo_feats.shape[0]
    rep = torch.cat([g[i][:, None, :].expand(B, n, g.shape[2]) for i, n in enumerate(np_cluster)], dim=1)  # :192-197
    h = torch.cat([lo_feats, rep], dim=2)                                                      # :200
    M = h.shape[1]
    h = h.reshape(B * M, -1)
    h = torch.relu(_bn(_conv(h, sd, "conv_2"), sd, "bn_2", training, stats))                   # :203
    if keep_2 is not None:
        h = h * keep_2                                                                         # :204
    h = torch.relu(_bn(_conv(h, sd, "conv_3"), sd, "bn_3", training, stats))                   # :205
    if keep_3 is not None:
        h = h * keep_3                                                                         # :206
    h = _conv(h, sd, "conv_4")                                                                 # :207
    return h.reshape(B, M, -1).transpose(1, 2)


def forward_windows(sd_enc, sd_seg, x_windows, centroids, attn_mask=None, training=False, stats_enc=None,
                    stats_seg=None, taps=None):
    """The per-window encoder loop + attention head of train_pointnet-attention.py:396-435 /
    test_pointnet_att_segmen.py:160-177. x_windows: list of [B, N_w, 9]. Returns (logits, last F64).
    stats_enc / stats_seg (dicts): when given, BN running statistics in sd_* are updated in place, once per
    encoder call (9 sequential updates per step for W=9: quirk 7 of SURVEY 3.5).
    taps (dict): receives gl_feats [W, B, 256] / lo_feats [B, sumN, 64] (gradients retained) for the backward parity tests."""
    lo, gl, npc, F64 = [], [], [], None
    for xw in x_windows:
        out, F64 = base_pointnet(sd_enc, xw, training, stats_enc)
        lo.append(out[:, :, -64:])                                # train_...:411
        gl.append(out[:, 0, :-64])                                # :412
        npc.append(xw.shape[1])
    gl_feats = torch.stack(gl, 0)                                 # [W, B, 256]  (:417,432)
    lo_feats = torch.cat(lo, 1)
    if taps is not None:                                          # the two tensors that cross from the encoder to the head
        taps["gl_feats"], taps["lo_feats"] = gl_feats, lo_feats
        if gl_feats.requires_grad:
            gl_feats.retain_grad(); lo_feats.retain_grad()
    logits = seg_attention(sd_seg, gl_feats, lo_feats, centroids, npc, attn_mask, training, stats_seg)
    return logits, F64


def train_step_loss(logits, targets, F64, class_weights=(1., 2., 2., 1., 1.)):
    """CE(weight, ignore_index=-1, mean) + 0.001 * ||I - F F^T||   (train_pointnet-attention.py:138,445,463-466)"""
    w = torch.tensor(class_weights, dtype=logits.dtype)
    lp = torch.log_softmax(logits, dim=1)                          # [B, C, M]
    valid = targets != -1
    t = targets.clamp(min=0)
    picked = lp.gather(1, t[:, None, :]).squeeze(1)
    wt = w[t] * valid
    ce = -(picked * wt).sum() / wt.sum()
    eye = torch.eye(F64.shape[-1], dtype=F64.dtype)
    reg = torch.linalg.norm((eye - torch.bmm(F64, F64.transpose(2, 1))).reshape(-1))
    return ce + 0.001 * reg, ce, reg

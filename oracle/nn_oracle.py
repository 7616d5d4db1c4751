"""CPU ORACLE (test infrastructure, NOT product code) for the aggregation modules of the hot path.

numpy float32 restatements that follow the reference's operation order (reference =
tgm-team/tgm @ 5183dc9), one function per module:

* `temporal_attention`  tgm/nn/modules/attention.py:58-128 (eval mode: dropout = identity)
* `merge_layer`         tgm/nn/encoder/tgat.py:34-38
* `tgat_forward`        tgm/nn/encoder/tgat.py:122-149
* Time2Vec lives in oracle/recency_oracle.py::time2vec (float64 cosine of the float32 argument)

Parity pinning: tests/test_oracle_golden.py checks these against tests/golden/nn_*.npz, produced
by running the unmodified reference modules in the build container
(tests/golden/make_golden_nn.py), at 2e-6.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from oracle.recency_oracle import PADDED_NODE_ID, time2vec

f32 = np.float32


def _linear(x: np.ndarray, W: np.ndarray, b=None) -> np.ndarray:
    y = x.astype(f32) @ W.astype(f32).T
    return y if b is None else y + b.astype(f32)


def _t2v(p: Dict[str, np.ndarray], prefix: str, dt: np.ndarray) -> np.ndarray:
    w = p[prefix + 'w.weight'].reshape(-1)
    b = p[prefix + 'w.bias']
    flat = time2vec(np.asarray(dt).reshape(-1), w, b, fused=True).astype(f32)
    return flat.reshape(*np.asarray(dt).shape, -1)


def layer_norm(x: np.ndarray, w: np.ndarray, b: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    mu = x.mean(-1, keepdims=True, dtype=f32)
    var = ((x - mu) ** 2).mean(-1, keepdims=True, dtype=f32)
    return ((x - mu) / np.sqrt(var + f32(eps)) * w + b).astype(f32)


def temporal_attention(p: Dict[str, np.ndarray], prefix: str, n_heads: int, node_x, time_feat,
                       edge_feat, nbr_node_feat, nbr_time_feat, valid_nbr_mask) -> np.ndarray:
    """attention.py:93-128."""
    out_dim = p[prefix + 'W_Q.weight'].shape[0]
    S, k = valid_nbr_mask.shape
    pad = out_dim - node_x.shape[1] - time_feat.shape[1]
    X = np.pad(node_x, ((0, 0), (0, pad))) if pad else node_x  # :93
    R = np.concatenate([X, time_feat], 1).astype(f32)  # :95
    Q = _linear(R, p[prefix + 'W_Q.weight'])  # :96
    Z = np.concatenate([nbr_node_feat, edge_feat, nbr_time_feat], -1).astype(f32)  # :98
    Z = _linear(Z.reshape(S * k, -1), p[prefix + 'W_KV.weight']).reshape(S, k, -1)  # :99
    K, V = Z[:, :, :out_dim], Z[:, :, out_dim:]  # :100-101
    hd = out_dim // n_heads
    Qh = Q.reshape(S, n_heads, hd)
    Kh = K.reshape(S, k, n_heads, hd).transpose(0, 2, 1, 3)  # (S,H,k,hd) :103-105
    Vh = V.reshape(S, k, n_heads, hd).transpose(0, 2, 1, 3)
    A = np.einsum('shd,shnd->shn', Qh, Kh).astype(f32) * f32(hd ** -0.5)  # :108-109
    A = np.where(valid_nbr_mask[:, None, :], A, f32(-1e10))  # :117
    A = A - A.max(-1, keepdims=True)
    A = np.exp(A, dtype=f32)
    A = A / A.sum(-1, keepdims=True, dtype=f32)  # :118
    O = np.einsum('shn,shnd->shd', A, Vh).astype(f32).reshape(S, out_dim)  # :121-122
    out = _linear(O, p[prefix + 'W_O.weight'], p[prefix + 'W_O.bias'])  # :125
    return layer_norm(out + R, p[prefix + 'layer_norm.weight'], p[prefix + 'layer_norm.bias'])  # :127


def merge_layer(p: Dict[str, np.ndarray], prefix: str, x1, x2) -> np.ndarray:
    """tgat.py:34-38."""
    h = _linear(np.concatenate([x1, x2], 1), p[prefix + 'fc1.weight'], p[prefix + 'fc1.bias'])
    return _linear(np.maximum(h, 0), p[prefix + 'fc2.weight'], p[prefix + 'fc2.bias'])


def tgat_forward(p: Dict[str, np.ndarray], num_layers: int, n_heads: int, node_x: np.ndarray,
                 seed_nids: List[np.ndarray], seed_times: List[np.ndarray],
                 nbr_nids: List[np.ndarray], nbr_edge_x: List[np.ndarray],
                 nbr_edge_time: List[np.ndarray]) -> np.ndarray:
    """tgat.py:122-149.  `p` is the module's state_dict as numpy arrays."""
    z = {j: {} for j in range(num_layers + 1)}
    z[0][0] = node_x[seed_nids[0]]  # :131 (negative ids index from the end, as torch does)
    for i in range(1, num_layers + 1):
        z[0][i] = node_x[nbr_nids[i - 1].reshape(-1)]  # :132-134
    for j in range(1, num_layers + 1):
        for i in range(num_layers - j + 1):
            n = z[j - 1][i].shape[0]
            k = nbr_nids[j - 1].shape[-1]
            out = temporal_attention(
                p, f'attn.{j - 1}.', n_heads,
                node_x=z[j - 1][i],
                time_feat=_t2v(p, 'time_encoder.', np.zeros(n, np.int64)),
                nbr_node_feat=z[j - 1][i + 1].reshape(n, k, -1),
                edge_feat=nbr_edge_x[i],
                valid_nbr_mask=nbr_nids[i] != PADDED_NODE_ID,
                nbr_time_feat=_t2v(p, 'time_encoder.', seed_times[i][:, None] - nbr_edge_time[i]))
            z[j][i] = merge_layer(p, f'merge_layers.{j - 1}.', out, z[0][i])  # :148
    return z[num_layers][0]


# ---- DyGFormer (tgm/nn/encoder/dygformer.py) ------------------------------------------------------
def _gelu(x: np.ndarray) -> np.ndarray:
    """F.gelu (exact, erf form) in float64, rounded once."""
    try:
        from scipy.special import erf as _erf
        v = _erf(x.astype(np.float64) / np.sqrt(2.0))
    except ImportError:
        from math import erf
        v = np.vectorize(erf, otypes=[np.float64])(x.astype(np.float64) / np.sqrt(2.0))
    return (x * (0.5 * (1.0 + v))).astype(f32)


def cooccurrence_freq(src_nbrs: np.ndarray, dst_nbrs: np.ndarray):
    """NeighborCooccurrenceEncoder._count_nodes_freq (dygformer.py:33-51): per position the
    number of occurrences of that id in its own sequence and in the other one; 0 for padding."""
    B, L = src_nbrs.shape
    src_freq = np.zeros((B, L, 2), f32)
    dst_freq = np.zeros((B, L, 2), f32)
    for b in range(B):
        s, d = src_nbrs[b], dst_nbrs[b]
        for j in range(L):
            src_freq[b, j] = ((s == s[j]).sum(), (d == s[j]).sum())
            dst_freq[b, j] = ((d == d[j]).sum(), (s == d[j]).sum())
    src_freq[src_nbrs == PADDED_NODE_ID] = 0
    dst_freq[dst_nbrs == PADDED_NODE_ID] = 0
    return src_freq, dst_freq


def _cooc_encode(p, freq):
    pre = 'co_occurrence_encoder.neighbor_co_occurrence_encoder.'
    h = np.maximum(_linear(freq[..., None], p[pre + '0.weight'], p[pre + '0.bias']), 0)  # :68-70
    return _linear(h, p[pre + '2.weight'], p[pre + '2.bias']).sum(2).astype(f32)


def _mha(p, pre, x, num_heads):
    """nn.MultiheadAttention(batch_first=False) self-attention on x [B, T, E] (eval mode)."""
    B, T, E = x.shape
    hd = E // num_heads
    qkv = _linear(x, p[pre + 'in_proj_weight'], p[pre + 'in_proj_bias'])
    q, k, v = (qkv[..., i * E:(i + 1) * E].reshape(B, T, num_heads, hd).transpose(0, 2, 1, 3)
               for i in range(3))
    a = np.einsum('bhqd,bhkd->bhqk', q * f32(hd ** -0.5), k).astype(f32)
    a = np.exp(a - a.max(-1, keepdims=True), dtype=f32)
    a = a / a.sum(-1, keepdims=True, dtype=f32)
    o = np.einsum('bhqk,bhkd->bhqd', a, v).astype(f32).transpose(0, 2, 1, 3).reshape(B, T, E)
    return _linear(o, p[pre + 'out_proj.weight'], p[pre + 'out_proj.bias'])


def _transformer_layer(p, pre, x, num_heads):
    """TransformerEncoder.forward (dygformer.py:117-143), pre-LN, eval mode."""
    h = layer_norm(x, p[pre + 'norm_layers.0.weight'], p[pre + 'norm_layers.0.bias'])
    out = x + _mha(p, pre + 'multi_head_attention.', h, num_heads)
    h = layer_norm(out, p[pre + 'norm_layers.1.weight'], p[pre + 'norm_layers.1.bias'])
    h = _gelu(_linear(h, p[pre + 'linear_layers.0.weight'], p[pre + 'linear_layers.0.bias']))
    return out + _linear(h, p[pre + 'linear_layers.1.weight'], p[pre + 'linear_layers.1.bias'])


def dygformer_forward(p: Dict[str, np.ndarray], patch_size: int, num_layers: int, num_heads: int,
                      node_x, edge_index, edge_time, neighbours, neighbours_time,
                      neighbours_edge_feat):
    """DyGFormer.forward (dygformer.py:243-431); `p` is the module's state_dict as numpy."""
    src, dst = edge_index[0], edge_index[1]
    B = len(src)
    seqs = []
    for ids, sl in ((src, slice(0, B)), (dst, slice(B, 2 * B))):
        nb = np.concatenate([ids[:, None], neighbours[sl]], 1)  # :274-275
        nt = np.concatenate([edge_time[:, None], neighbours_time[sl]], 1)
        ef = np.concatenate([np.zeros((B, 1, neighbours_edge_feat.shape[2]), f32),
                             neighbours_edge_feat[sl]], 1)
        nf = node_x[nb].astype(f32)
        nf[nb == PADDED_NODE_ID] = 0  # :298-299
        tf = _t2v(p, 'time_encoder.', edge_time[:, None] - nt)
        tf[nb == PADDED_NODE_ID] = 0  # :308-310
        seqs.append((nb, nf, ef, tf))
    cs, cd = cooccurrence_freq(seqs[0][0], seqs[1][0])
    cooc = (_cooc_encode(p, cs), _cooc_encode(p, cd))
    L = seqs[0][0].shape[1]
    NP = L // patch_size
    tokens = []
    for side in range(2):
        _, nf, ef, tf = seqs[side]
        chans = []
        for name, feat in (('node', nf), ('edge', ef), ('time', tf),
                           ('neighbor_co_occurrence', cooc[side])):
            patches = feat.reshape(B, NP, patch_size * feat.shape[2])  # _get_patches :433-444
            chans.append(_linear(patches, p[f'projection_layer.{name}.weight'],
                                 p[f'projection_layer.{name}.bias']))
        tokens.append(np.stack(chans, 2).reshape(B, NP, -1))  # :401-413
    x = np.concatenate(tokens, 1).astype(f32)
    for i in range(num_layers):
        x = _transformer_layer(p, f'transformers.{i}.', x, num_heads)
    outs = []
    for side in range(2):
        pooled = x[:, side * NP:(side + 1) * NP].mean(1, dtype=f32)  # :424-425
        outs.append(_linear(pooled, p['output_layer.weight'], p['output_layer.bias']))
    return outs[0], outs[1]


# ---- DyGFormer backward (checker for tgm_dyg_backward) ----------------------------------------------
def dygformer_backward(p: Dict[str, np.ndarray], patch_size: int, num_layers: int, num_heads: int,
                       node_x, edge_index, edge_time, neighbours, neighbours_time,
                       neighbours_edge_feat, G_src, G_dst) -> Dict[str, np.ndarray]:
    """Gradients of sum(z_src * G_src) + sum(z_dst * G_dst) w.r.t. every parameter of
    DyGFormer.forward (dygformer.py:243-431, eval mode), by hand in float64 -- the chain rule of
    the forward above, pinned against the reference's autograd (tests/golden/nn_dyggrad_*.npz)."""
    f64 = np.float64
    P = {k: v.astype(f64) for k, v in p.items()}
    src, dst = edge_index[0], edge_index[1]
    B = len(src)
    w_t, b_t = P['time_encoder.w.weight'].reshape(-1), P['time_encoder.w.bias']
    pre_c = 'co_occurrence_encoder.neighbor_co_occurrence_encoder.'
    w1, b1 = P[pre_c + '0.weight'].reshape(-1), P[pre_c + '0.bias']
    W2, b2 = P[pre_c + '2.weight'], P[pre_c + '2.bias']
    seqs = []
    for ids, sl in ((src, slice(0, B)), (dst, slice(B, 2 * B))):
        nb = np.concatenate([ids[:, None], neighbours[sl]], 1)
        nt = np.concatenate([edge_time[:, None], neighbours_time[sl]], 1)
        ef = np.concatenate([np.zeros((B, 1, neighbours_edge_feat.shape[2])),
                             neighbours_edge_feat[sl].astype(f64)], 1)
        nf = node_x[nb].astype(f64)
        pad = nb == PADDED_NODE_ID
        nf[pad] = 0
        dt = (edge_time[:, None] - nt).astype(np.float32).astype(f64)
        arg = dt[..., None] * w_t + b_t
        tf = np.cos(arg)
        tf[pad] = 0
        seqs.append(dict(nb=nb, nf=nf, ef=ef, tf=tf, arg=arg, dt=dt, pad=pad))
    cs, cd = cooccurrence_freq(seqs[0]['nb'], seqs[1]['nb'])
    for sd, freq in zip(seqs, (cs, cd)):
        freq = freq.astype(f64)                               # [B, L, 2]
        hpre = freq[..., None] * w1 + b1                      # [B, L, 2, C]
        h = np.maximum(hpre, 0)
        sd.update(freq=freq, hpre=hpre, h=h, cooc=(h @ W2.T + b2).sum(2))
    L = seqs[0]['nb'].shape[1]
    NP = L // patch_size
    names = ('node', 'edge', 'time', 'neighbor_co_occurrence')
    C = P['projection_layer.node.weight'].shape[0]
    tokens = []
    for sd in seqs:
        chans = []
        sd['patches'] = []
        for name, feat in zip(names, (sd['nf'], sd['ef'], sd['tf'], sd['cooc'])):
            pt = feat.reshape(B, NP, patch_size * feat.shape[2])
            sd['patches'].append(pt)
            chans.append(pt @ P[f'projection_layer.{name}.weight'].T + P[f'projection_layer.{name}.bias'])
        tokens.append(np.stack(chans, 2).reshape(B, NP, -1))
    x = np.concatenate(tokens, 1)                             # [B, T, E]
    T, E = x.shape[1], x.shape[2]
    H, hd = num_heads, x.shape[2] // num_heads

    def ln_fwd(v, w, b, eps=1e-5):
        mu = v.mean(-1, keepdims=True)
        var = ((v - mu) ** 2).mean(-1, keepdims=True)
        rstd = 1.0 / np.sqrt(var + eps)
        xh = (v - mu) * rstd
        return xh * w + b, xh, rstd

    def ln_bwd(dy, xh, rstd, w):
        dxh = dy * w
        return rstd * (dxh - dxh.mean(-1, keepdims=True) - xh * (dxh * xh).mean(-1, keepdims=True))

    from math import sqrt, pi
    try:
        from scipy.special import erf as _erf
    except ImportError:  # pragma: no cover
        from math import erf
        _erf = np.vectorize(erf, otypes=[np.float64])
    cache = []
    for i in range(num_layers):
        pre = f'transformers.{i}.'
        xn0, xh0, r0 = ln_fwd(x, P[pre + 'norm_layers.0.weight'], P[pre + 'norm_layers.0.bias'])
        qkv = xn0 @ P[pre + 'multi_head_attention.in_proj_weight'].T + P[pre + 'multi_head_attention.in_proj_bias']
        q, k, v = (qkv[..., j * E:(j + 1) * E].reshape(B, T, H, hd).transpose(0, 2, 1, 3) for j in range(3))
        s = np.einsum('bhqd,bhkd->bhqk', q, k) * hd ** -0.5
        a = np.exp(s - s.max(-1, keepdims=True))
        a /= a.sum(-1, keepdims=True)
        o = np.einsum('bhqk,bhkd->bhqd', a, v).transpose(0, 2, 1, 3).reshape(B, T, E)
        x1 = x + o @ P[pre + 'multi_head_attention.out_proj.weight'].T + P[pre + 'multi_head_attention.out_proj.bias']
        xn1, xh1, r1 = ln_fwd(x1, P[pre + 'norm_layers.1.weight'], P[pre + 'norm_layers.1.bias'])
        f1 = xn1 @ P[pre + 'linear_layers.0.weight'].T + P[pre + 'linear_layers.0.bias']
        g1 = f1 * 0.5 * (1 + _erf(f1 / sqrt(2)))
        x2 = x1 + g1 @ P[pre + 'linear_layers.1.weight'].T + P[pre + 'linear_layers.1.bias']
        cache.append(dict(xn0=xn0, xh0=xh0, r0=r0, q=q, k=k, v=v, a=a, o=o, xn1=xn1, xh1=xh1, r1=r1,
                          f1=f1, g1=g1))
        x = x2
    g: Dict[str, np.ndarray] = {}
    dx = np.zeros_like(x)
    Wo = P['output_layer.weight']
    g['output_layer.weight'] = np.zeros_like(Wo)
    g['output_layer.bias'] = np.zeros_like(P['output_layer.bias'])
    for side, G in enumerate((G_src.astype(f64), G_dst.astype(f64))):
        pooled = x[:, side * NP:(side + 1) * NP].mean(1)
        g['output_layer.weight'] += G.T @ pooled
        g['output_layer.bias'] += G.sum(0)
        dx[:, side * NP:(side + 1) * NP] = (G @ Wo)[:, None, :] / NP
    for i in reversed(range(num_layers)):
        pre, c = f'transformers.{i}.', cache[i]
        W2f, W1f = P[pre + 'linear_layers.1.weight'], P[pre + 'linear_layers.0.weight']
        g[pre + 'linear_layers.1.weight'] = np.einsum('bte,btf->ef', dx, c['g1'])
        g[pre + 'linear_layers.1.bias'] = dx.sum((0, 1))
        dg1 = dx @ W2f
        df1 = dg1 * (0.5 * (1 + _erf(c['f1'] / sqrt(2))) + c['f1'] * np.exp(-c['f1'] ** 2 / 2) / sqrt(2 * pi))
        g[pre + 'linear_layers.0.weight'] = np.einsum('btf,bte->fe', df1, c['xn1'])
        g[pre + 'linear_layers.0.bias'] = df1.sum((0, 1))
        dxn1 = df1 @ W1f
        g[pre + 'norm_layers.1.weight'] = (dxn1 * c['xh1']).sum((0, 1))
        g[pre + 'norm_layers.1.bias'] = dxn1.sum((0, 1))
        dx1 = dx + ln_bwd(dxn1, c['xh1'], c['r1'], P[pre + 'norm_layers.1.weight'])
        Wout = P[pre + 'multi_head_attention.out_proj.weight']
        g[pre + 'multi_head_attention.out_proj.weight'] = np.einsum('bte,btf->ef', dx1, c['o'])
        g[pre + 'multi_head_attention.out_proj.bias'] = dx1.sum((0, 1))
        do = (dx1 @ Wout).reshape(B, T, H, hd).transpose(0, 2, 1, 3)
        da = np.einsum('bhqd,bhkd->bhqk', do, c['v'])
        dv = np.einsum('bhqk,bhqd->bhkd', c['a'], do)
        ds = c['a'] * (da - (da * c['a']).sum(-1, keepdims=True)) * hd ** -0.5
        dq = np.einsum('bhqk,bhkd->bhqd', ds, c['k'])
        dk = np.einsum('bhqk,bhqd->bhkd', ds, c['q'])
        dqkv = np.concatenate([t_.transpose(0, 2, 1, 3).reshape(B, T, E) for t_ in (dq, dk, dv)], -1)
        Win = P[pre + 'multi_head_attention.in_proj_weight']
        g[pre + 'multi_head_attention.in_proj_weight'] = np.einsum('btf,bte->fe', dqkv, c['xn0'])
        g[pre + 'multi_head_attention.in_proj_bias'] = dqkv.sum((0, 1))
        dxn0 = dqkv @ Win
        g[pre + 'norm_layers.0.weight'] = (dxn0 * c['xh0']).sum((0, 1))
        g[pre + 'norm_layers.0.bias'] = dxn0.sum((0, 1))
        dx = dx1 + ln_bwd(dxn0, c['xh0'], c['r0'], P[pre + 'norm_layers.0.weight'])
    # projections and the two learnable feature channels
    g['time_encoder.w.weight'] = np.zeros_like(w_t)
    g['time_encoder.w.bias'] = np.zeros_like(b_t)
    for key_, like in ((pre_c + '0.weight', w1), (pre_c + '0.bias', b1), (pre_c + '2.weight', W2),
                       (pre_c + '2.bias', b2)):
        g[key_] = np.zeros_like(like)
    for name in names:
        g[f'projection_layer.{name}.weight'] = np.zeros_like(P[f'projection_layer.{name}.weight'])
        g[f'projection_layer.{name}.bias'] = np.zeros_like(P[f'projection_layer.{name}.bias'])
    for side, sd in enumerate(seqs):
        dtok = dx[:, side * NP:(side + 1) * NP].reshape(B, NP, 4, C)
        for ci, name in enumerate(names):
            dch = dtok[:, :, ci]                                              # [B, NP, C]
            g[f'projection_layer.{name}.weight'] += np.einsum('bpc,bpk->ck', dch, sd['patches'][ci])
            g[f'projection_layer.{name}.bias'] += dch.sum((0, 1))
            if ci < 2:
                continue
            dfeat = (dch @ P[f'projection_layer.{name}.weight']).reshape(B, L, -1)  # un-patch
            if ci == 2:
                dfeat = np.where(sd['pad'][..., None], 0.0, dfeat)
                darg = -np.sin(sd['arg']) * dfeat
                g['time_encoder.w.weight'] += (darg * sd['dt'][..., None]).sum((0, 1))
                g['time_encoder.w.bias'] += darg.sum((0, 1))
            else:
                g[pre_c + '2.bias'] += 2 * dfeat.sum((0, 1))
                g[pre_c + '2.weight'] += np.einsum('blc,blj->cj', dfeat, sd['h'].sum(2))
                dh = (dfeat @ W2)[:, :, None, :] * (sd['hpre'] > 0)             # [B, L, 2, C]
                g[pre_c + '0.weight'] += (dh * sd['freq'][..., None]).sum((0, 1, 2))
                g[pre_c + '0.bias'] += dh.sum((0, 1, 2))
    g['time_encoder.w.weight'] = g['time_encoder.w.weight'].reshape(-1, 1)
    g[pre_c + '0.weight'] = g[pre_c + '0.weight'].reshape(-1, 1)
    return g

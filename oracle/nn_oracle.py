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

"""EAGER-PYTORCH restatements of the hot path (test infrastructure, NOT product code).

The reference (tgm-team/tgm @ 5183dc9) is eager PyTorch: on a GPU box its own `device='cuda'`
mode is these very tensor ops launched one by one.  The reference tree cannot travel to the GPU
box, so the benches time THIS module on the B200 beside every row as `eager_cuda_baseline`: the
same operation order, the same intermediate tensors, the same host syncs as the reference, written
with plain torch ops on whatever device the inputs live on.  Only `tests/`, `bench*.py` baseline
legs and `__graft_entry__.smoke()` may import it; nothing under `tgm_b200/` does.

  TorchRing            tgm/hooks/neighbors/recency.py:93-97, 111-117, 239-321, 323-399
  time2vec             tgm/nn/modules/time_encoding.py:22-24
  temporal_attention   tgm/nn/modules/attention.py:58-128 (eval: dropout = identity)
  merge_layer          tgm/nn/encoder/tgat.py:34-38
  tgat_forward         tgm/nn/encoder/tgat.py:122-149
  dygformer_forward    tgm/nn/encoder/dygformer.py:243-431 (+ :13-77, :80-143, :433-444)
  TorchTGNMemory       tgm/nn/encoder/tgn.py:80-251 (IdentityMessage + LastAggregator, with the
                       reference's per-node Python dict message store :183-184, :218-243)

Pinned on CPU against the numpy oracles and the reference-generated fixtures
(tests/test_torch_eager.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

PADDED_NODE_ID = -1


# ---- sampler -------------------------------------------------------------------------------------
class TorchRing:
    """RecencyNeighborHook's state machine with the reference's op structure."""

    def __init__(self, num_nodes: int, num_nbrs: Sequence[int], edge_x_dim: int = 0,
                 directed: bool = False, device='cpu') -> None:
        self.N, self.num_nbrs, self.B = int(num_nodes), list(num_nbrs), max(num_nbrs)
        self.D, self.directed, self.device = int(edge_x_dim), directed, torch.device(device)
        self.reset_state()

    def reset_state(self) -> None:  # recency.py:111-117
        dev = self.device
        self.ids = torch.full((self.N, self.B), PADDED_NODE_ID, dtype=torch.int32, device=dev)
        self.times = torch.zeros((self.N, self.B), dtype=torch.int64, device=dev)
        self.feats = torch.zeros((self.N, self.B, self.D), dtype=torch.float32, device=dev)
        self.write_pos = torch.zeros(self.N, dtype=torch.int32, device=dev)

    def query(self, seeds: Tensor, tq: Tensor, k: int) -> Tuple[Tensor, Tensor, Tensor]:
        B = self.B
        # the reference warns when a query precedes every stored time: a reduction over the whole
        # (N, B) buffer and a host sync on every call (recency.py:242-251)
        if bool(tq.min() < self.times.min()):
            pass
        rows = seeds.long()
        ring_ids, ring_t, ring_x = self.ids[rows], self.times[rows], self.feats[rows]  # :256-259
        wp = self.write_pos[rows].long()
        ar = torch.arange(B, device=seeds.device)
        slot = (wp[:, None] + ar[None, :]) % B                      # oldest .. newest (:263-264)
        t_un = torch.gather(ring_t, 1, slot)
        id_un = torch.gather(ring_ids, 1, slot)
        ok = (t_un < tq[:, None]) & (id_un != PADDED_NODE_ID)        # :267-269
        last = torch.where(ok.any(1), (ok * ar).max(1).values, torch.full_like(wp, -1))
        col = last[:, None] - torch.arange(k - 1, -1, -1, device=seeds.device)[None, :]
        valid = col >= 0
        src_slot = torch.gather(slot, 1, col.clamp(min=0))
        out_id = torch.where(valid, torch.gather(ring_ids, 1, src_slot),
                             torch.full_like(src_slot, PADDED_NODE_ID, dtype=torch.int32))
        out_t = torch.where(valid, torch.gather(ring_t, 1, src_slot), torch.zeros_like(src_slot))
        out_x = torch.gather(ring_x, 1, src_slot[:, :, None].expand(-1, -1, self.D))
        out_x = torch.where(valid[:, :, None], out_x, torch.zeros((), device=seeds.device))
        return out_id, out_t, out_x                                  # :287-319

    def update(self, src: Tensor, dst: Tensor, t: Tensor, x: Optional[Tensor]) -> None:
        n = src.numel()
        if x is None:
            x = torch.zeros((n, self.D), device=src.device)          # :325-329
        if self.directed:
            node, nbr, tt, xx = src, dst, t, x
        else:
            node, nbr = torch.cat([src, dst]), torch.cat([dst, src])  # :339-342
            tt, xx = torch.cat([t, t]), torch.cat([x, x])
        key = node.long() * (tt.max() + 1) + tt                      # :347-348 (ideal: int64 key)
        perm = torch.argsort(key, stable=True)                       # :349
        node, nbr, tt, xx = node[perm], nbr[perm], tt[perm], xx[perm]
        uniq, counts = torch.unique_consecutive(node, return_counts=True)  # :364
        starts = torch.cumsum(counts, 0) - counts
        run = torch.repeat_interleave(torch.arange(uniq.numel(), device=node.device), counts)
        pos = torch.arange(node.numel(), device=node.device) - starts[run]
        keep = pos >= (counts[run] - self.B)                          # last B per node (:373)
        node, nbr, tt, xx, run, pos = node[keep], nbr[keep], tt[keep], xx[keep], run[keep], pos[keep]
        off = pos - torch.clamp(counts[run] - self.B, min=0)
        widx = (self.write_pos[node.long()].long() + off) % self.B    # :389
        self.ids[node.long(), widx] = nbr                             # :392-394
        self.times[node.long(), widx] = tt
        self.feats[node.long(), widx] = xx
        self.write_pos.index_add_(0, node.long(), torch.ones_like(node, dtype=torch.int32))  # :397-399

    def hook_call(self, seeds: Tensor, tq: Tensor, src: Tensor, dst: Tensor, t: Tensor,
                  x: Optional[Tensor]):
        out = []
        if seeds.numel():
            s, q = seeds, tq
            for hop, k in enumerate(self.num_nbrs):
                if hop > 0:
                    s, q = out[-1][2].reshape(-1), out[-1][3].reshape(-1)
                nid, nt, nx = self.query(s, q, k)
                out.append((s, q, nid, nt, nx))
            if src.numel():
                self.update(src, dst, t, x)
        return out


# ---- aggregation ---------------------------------------------------------------------------------
def time2vec(p: Dict[str, Tensor], prefix: str, dt: Tensor) -> Tensor:
    """cos(Linear(1, d)(dt.float()))  (time_encoding.py:22-24)."""
    return torch.cos(F.linear(dt.float().unsqueeze(-1), p[prefix + 'w.weight'], p[prefix + 'w.bias']))


def temporal_attention(p: Dict[str, Tensor], prefix: str, n_heads: int, node_x, time_feat,
                       edge_feat, nbr_node_feat, nbr_time_feat, valid_nbr_mask) -> Tensor:
    out_dim = p[prefix + 'W_Q.weight'].shape[0]
    S, k = valid_nbr_mask.shape
    pad = out_dim - node_x.shape[1] - time_feat.shape[1]
    if pad:
        node_x = F.pad(node_x, (0, pad))                               # attention.py:93
    R = torch.cat([node_x, time_feat], 1)                             # :95
    Q = F.linear(R, p[prefix + 'W_Q.weight'])                         # :96
    Z = torch.cat([nbr_node_feat, edge_feat, nbr_time_feat], -1)      # :98
    Z = F.linear(Z, p[prefix + 'W_KV.weight'])                        # :99
    K, V = Z[:, :, :out_dim], Z[:, :, out_dim:]
    hd = out_dim // n_heads
    Qh = Q.reshape(S, n_heads, hd)
    Kh = K.reshape(S, k, n_heads, hd).permute(0, 2, 1, 3)
    Vh = V.reshape(S, k, n_heads, hd).permute(0, 2, 1, 3)
    A = torch.einsum('shd,shnd->shn', Qh, Kh) * hd ** -0.5            # :108-109
    A = A.masked_fill(~valid_nbr_mask[:, None, :], -1e10)             # :117
    A = torch.softmax(A, -1)                                          # :118
    O = torch.einsum('shn,shnd->shd', A, Vh).reshape(S, out_dim)      # :121-122
    out = F.linear(O, p[prefix + 'W_O.weight'], p[prefix + 'W_O.bias'])
    return F.layer_norm(out + R, (out_dim,), p[prefix + 'layer_norm.weight'],
                        p[prefix + 'layer_norm.bias'])               # :127


def merge_layer(p: Dict[str, Tensor], prefix: str, x1: Tensor, x2: Tensor) -> Tensor:
    h = F.linear(torch.cat([x1, x2], 1), p[prefix + 'fc1.weight'], p[prefix + 'fc1.bias'])
    return F.linear(torch.relu(h), p[prefix + 'fc2.weight'], p[prefix + 'fc2.bias'])


def tgat_forward(p: Dict[str, Tensor], num_layers: int, n_heads: int, node_x: Tensor,
                 seed_nids: List[Tensor], seed_times: List[Tensor], nbr_nids: List[Tensor],
                 nbr_edge_x: List[Tensor], nbr_edge_time: List[Tensor]) -> Tensor:
    z = {j: {} for j in range(num_layers + 1)}
    z[0][0] = node_x[seed_nids[0].long()]                             # tgat.py:131
    for i in range(1, num_layers + 1):
        z[0][i] = node_x[nbr_nids[i - 1].reshape(-1).long()]          # :132-134
    for j in range(1, num_layers + 1):
        for i in range(num_layers - j + 1):
            n = z[j - 1][i].shape[0]
            k = nbr_nids[j - 1].shape[-1]
            out = temporal_attention(
                p, f'attn.{j - 1}.', n_heads, node_x=z[j - 1][i],
                time_feat=time2vec(p, 'time_encoder.', torch.zeros(n, device=node_x.device)),
                nbr_node_feat=z[j - 1][i + 1].reshape(n, k, -1), edge_feat=nbr_edge_x[i],
                valid_nbr_mask=nbr_nids[i] != PADDED_NODE_ID,
                nbr_time_feat=time2vec(p, 'time_encoder.',
                                       seed_times[i][:, None] - nbr_edge_time[i]))
            z[j][i] = merge_layer(p, f'merge_layers.{j - 1}.', out, z[0][i])  # :148
    return z[num_layers][0]


def _cooccurrence(src_nbrs: Tensor, dst_nbrs: Tensor):
    """dygformer.py:33-51 with broadcast compares (B, L, L)."""
    def count(a, b):
        own = (a[:, :, None] == a[:, None, :]).sum(-1)
        other = (a[:, :, None] == b[:, None, :]).sum(-1)
        f = torch.stack([own, other], -1).float()
        return f.masked_fill((a == PADDED_NODE_ID)[:, :, None], 0.0)
    return count(src_nbrs, dst_nbrs), count(dst_nbrs, src_nbrs)


def dygformer_forward(p: Dict[str, Tensor], patch_size: int, num_layers: int, num_heads: int,
                      node_x, edge_index, edge_time, neighbours, neighbours_time,
                      neighbours_edge_feat):
    src, dst = edge_index[0], edge_index[1]
    B = src.numel()
    seqs = []
    for ids, sl in ((src, slice(0, B)), (dst, slice(B, 2 * B))):
        nb = torch.cat([ids[:, None], neighbours[sl]], 1)                     # :274-275
        nt = torch.cat([edge_time[:, None], neighbours_time[sl]], 1)
        ef = torch.cat([torch.zeros((B, 1, neighbours_edge_feat.shape[2]), device=src.device),
                        neighbours_edge_feat[sl]], 1)
        padm = nb == PADDED_NODE_ID
        nf = node_x[nb.long()].masked_fill(padm[:, :, None], 0.0)             # :298-299
        tf = time2vec(p, 'time_encoder.', edge_time[:, None] - nt).masked_fill(padm[:, :, None], 0.0)
        seqs.append((nb, nf, ef, tf))
    pre = 'co_occurrence_encoder.neighbor_co_occurrence_encoder.'
    cooc = []
    for f in _cooccurrence(seqs[0][0], seqs[1][0]):
        h = torch.relu(F.linear(f[..., None], p[pre + '0.weight'], p[pre + '0.bias']))  # :68-70
        cooc.append(F.linear(h, p[pre + '2.weight'], p[pre + '2.bias']).sum(2))
    L = seqs[0][0].shape[1]
    NP = L // patch_size
    tokens = []
    for side in range(2):
        _, nf, ef, tf = seqs[side]
        chans = []
        for name, feat in (('node', nf), ('edge', ef), ('time', tf),
                           ('neighbor_co_occurrence', cooc[side])):
            patches = feat.reshape(B, NP, patch_size * feat.shape[2])           # :433-444
            chans.append(F.linear(patches, p[f'projection_layer.{name}.weight'],
                                  p[f'projection_layer.{name}.bias']))
        tokens.append(torch.stack(chans, 2).reshape(B, NP, -1))               # :401-413
    x = torch.cat(tokens, 1)
    E = x.shape[2]
    hd = E // num_heads
    for i in range(num_layers):                                               # :117-143, pre-LN
        tp = f'transformers.{i}.'
        h = F.layer_norm(x, (E,), p[tp + 'norm_layers.0.weight'], p[tp + 'norm_layers.0.bias'])
        qkv = F.linear(h, p[tp + 'multi_head_attention.in_proj_weight'],
                       p[tp + 'multi_head_attention.in_proj_bias'])
        q, k_, v = (qkv[..., j * E:(j + 1) * E].reshape(B, -1, num_heads, hd).permute(0, 2, 1, 3)
                    for j in range(3))
        a = torch.softmax(torch.matmul(q * hd ** -0.5, k_.transpose(-1, -2)), -1)
        o = torch.matmul(a, v).permute(0, 2, 1, 3).reshape(B, -1, E)
        x = x + F.linear(o, p[tp + 'multi_head_attention.out_proj.weight'],
                         p[tp + 'multi_head_attention.out_proj.bias'])
        h = F.layer_norm(x, (E,), p[tp + 'norm_layers.1.weight'], p[tp + 'norm_layers.1.bias'])
        h = F.gelu(F.linear(h, p[tp + 'linear_layers.0.weight'], p[tp + 'linear_layers.0.bias']))
        x = x + F.linear(h, p[tp + 'linear_layers.1.weight'], p[tp + 'linear_layers.1.bias'])
    outs = []
    for side in range(2):
        pooled = x[:, side * NP:(side + 1) * NP].mean(1)                      # :424-425
        outs.append(F.linear(pooled, p['output_layer.weight'], p['output_layer.bias']))
    return outs[0], outs[1]


# ---- TGN memory ------------------------------------------------------------------------------------
class TorchTGNMemory:
    """TGNMemory (IdentityMessage + LastAggregator + GRUCell) with the reference's per-node Python
    dict message store (tgn.py:183-184, :218-243) -- the structure that makes it host-bound."""

    def __init__(self, num_nodes: int, raw_msg_dim: int, memory_dim: int, time_dim: int,
                 params: Dict[str, Tensor], device='cpu') -> None:
        self.N, self.D, self.M, self.TD = num_nodes, raw_msg_dim, memory_dim, time_dim
        self.p, self.device, self.training = params, torch.device(device), True
        self.reset_state()

    def reset_state(self) -> None:
        self.memory = torch.zeros((self.N, self.M), device=self.device)
        self.last_update = torch.zeros(self.N, dtype=torch.long, device=self.device)
        self._reset_message_store()

    def _reset_message_store(self) -> None:                                   # :180-185
        i = torch.empty(0, dtype=torch.long, device=self.device)
        msg = torch.empty((0, self.D), device=self.device)
        self.msg_s = {j: (i, i, i, msg) for j in range(self.N)}
        self.msg_d = {j: (i, i, i, msg) for j in range(self.N)}

    def forward(self, n_id: Tensor):
        if self.training:
            return self._get_updated_memory(n_id)
        return self.memory[n_id], self.last_update[n_id]

    def update_state(self, src: Tensor, dst: Tensor, t: Tensor, raw_msg: Tensor) -> None:
        n_id = torch.cat([src, dst]).unique()
        if self.training:
            self._update_memory(n_id)
            self._update_msg_store(src, dst, t, raw_msg, self.msg_s)
            self._update_msg_store(dst, src, t, raw_msg, self.msg_d)
        else:
            self._update_msg_store(src, dst, t, raw_msg, self.msg_s)
            self._update_msg_store(dst, src, t, raw_msg, self.msg_d)
            self._update_memory(n_id)

    def _update_memory(self, n_id: Tensor) -> None:
        mem, lu = self._get_updated_memory(n_id)
        self.memory[n_id], self.last_update[n_id] = mem, lu

    def _compute_msg(self, n_id: Tensor, store):                              # :231-243
        data = [store[i] for i in n_id.tolist()]
        src, dst, t, raw = (torch.cat(c, 0) for c in zip(*data))
        t_rel = t - self.last_update[src]
        t_enc = time2vec(self.p, 'time_enc.', t_rel)
        return torch.cat([self.memory[src], self.memory[dst], raw, t_enc], -1), t, src

    def _get_updated_memory(self, n_id: Tensor):                              # :192-216
        assoc = torch.empty(self.N, dtype=torch.long, device=self.device)
        assoc[n_id] = torch.arange(n_id.numel(), device=self.device)
        msg_s, t_s, src_s = self._compute_msg(n_id, self.msg_s)
        msg_d, t_d, src_d = self._compute_msg(n_id, self.msg_d)
        idx, msg, t = torch.cat([src_s, src_d]), torch.cat([msg_s, msg_d]), torch.cat([t_s, t_d])
        # LastAggregator (:43-56): per node the message with the largest time (first among ties)
        n = n_id.numel()
        aggr = msg.new_zeros((n, msg.shape[1]))
        if idx.numel():
            rows = assoc[idx]
            tf = t.float()
            best_t = torch.full((n,), float('-inf'), device=self.device).scatter_reduce(
                0, rows, tf, reduce='amax', include_self=True)
            is_best = tf == best_t[rows]
            pos = torch.arange(idx.numel(), device=self.device)
            first = torch.full((n,), idx.numel(), dtype=torch.long, device=self.device).scatter_reduce(
                0, rows[is_best], pos[is_best], reduce='amin', include_self=True)
            has = first < idx.numel()
            aggr[has] = msg[first[has]]
        gi = F.linear(aggr, self.p['memory_updater.weight_ih'], self.p['memory_updater.bias_ih'])
        gh = F.linear(self.memory[n_id], self.p['memory_updater.weight_hh'],
                      self.p['memory_updater.bias_hh'])
        M = self.M
        r = torch.sigmoid(gi[:, :M] + gh[:, :M])
        zg = torch.sigmoid(gi[:, M:2 * M] + gh[:, M:2 * M])
        ng = torch.tanh(gi[:, 2 * M:] + r * gh[:, 2 * M:])
        memory = (1 - zg) * ng + zg * self.memory[n_id]
        lu_all = torch.zeros(self.N, dtype=torch.long, device=self.device)
        if idx.numel():
            lu_all = lu_all.scatter_reduce(0, idx, t, reduce='amax', include_self=False)
        return memory, lu_all[n_id]

    def _update_msg_store(self, src, dst, t, raw_msg, store) -> None:         # :218-229
        n_id, perm = src.sort(stable=True)
        n_id, count = n_id.unique_consecutive(return_counts=True)
        for i, idx in zip(n_id.tolist(), perm.split(count.tolist())):
            store[i] = (src[idx], dst[idx], t[idx], raw_msg[idx])

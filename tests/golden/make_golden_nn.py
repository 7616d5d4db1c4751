"""Golden vectors for the aggregation modules, from the UNMODIFIED reference (tgm-team/tgm @
5183dc9) run in the build container:  python tests/golden/make_golden_nn.py

The reference's own nn tests are shape/NaN-only (test/unit/test_nn/test_temporal_attention.py:
25-62, test_tgat.py:6-46), so the fixtures are produced by the live reference modules in eval()
mode with seeded weights, on batches sampled by the reference's own RecencyNeighborHook.
Writes tests/golden/nn_*.npz (inputs, state_dict, outputs).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shim import import_reference  # noqa: E402

import_reference()
from tgm import DGraph  # noqa: E402
from tgm.data import DGData, DGDataLoader  # noqa: E402
from tgm.hooks import HookManager, RecencyNeighborHook  # noqa: E402
from tgm.nn import TGAT  # noqa: E402
from tgm.nn.modules import TemporalAttention, Time2Vec  # noqa: E402


def sampled_batch(N, E, T, D, bs, num_nbrs, which, seed):
    rng = np.random.default_rng(seed)
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    t = np.sort(rng.integers(0, T, E))
    x = rng.standard_normal((E, D)).astype(np.float32)
    data = DGData.from_raw(torch.from_numpy(t), torch.from_numpy(np.stack([src, dst], 1)).int(),
                           torch.from_numpy(x))
    dg = DGraph(data)
    hm = HookManager(keys=['g'])
    hm.register('g', RecencyNeighborHook(num_nodes=N, num_nbrs=list(num_nbrs),
                                         seed_nodes_keys=['edge_src', 'edge_dst'],
                                         seed_times_keys=['edge_time', 'edge_time']))
    with hm.activate('g'):
        for b, batch in enumerate(DGDataLoader(dg, batch_size=bs, hook_manager=hm)):
            if b == which:
                return batch
    raise RuntimeError('batch not reached')


def randomise(module, seed):
    """Generic (non-default) parameter values, incl. LayerNorm affine and the Time2Vec bias."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in module.named_parameters():
            if name.endswith('time_encoder.w.weight') or name == 'w.weight':
                continue  # keep the shipped geometric frequencies (time_encoding.py:18-19)
            if 'layer_norm.weight' in name:
                prm.copy_(1 + 0.3 * torch.randn(prm.shape, generator=g))
            elif prm.ndim == 1:
                prm.copy_(0.3 * torch.randn(prm.shape, generator=g))
            else:
                prm.copy_(torch.randn(prm.shape, generator=g) / prm.shape[1] ** 0.5)


def state_np(module, prefix='p.'):
    return {prefix + k: v.detach().numpy() for k, v in module.state_dict().items()}


def save_tgat(name, N, E, T, D, bs, num_nbrs, which, node_dim, time_dim, embed, heads, zero_bias):
    batch = sampled_batch(N, E, T, D, bs, num_nbrs, which, seed=len(name))
    torch.manual_seed(7)
    model = TGAT(node_dim=node_dim, edge_dim=D, time_dim=time_dim, embed_dim=embed,
                 num_layers=len(num_nbrs), n_heads=heads).eval()
    randomise(model, 11)
    if zero_bias:
        with torch.no_grad():
            model.time_encoder.w.bias.zero_()
    node_x = torch.randn(N, node_dim, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        out = model(node_x, batch.seed_nids, batch.seed_times, batch.nbr_nids, batch.nbr_edge_x,
                    batch.nbr_edge_time)
    d = dict(node_x=node_x.numpy(), out=out.numpy(), n_heads=np.int64(heads),
             num_layers=np.int64(len(num_nbrs)), **state_np(model))
    for h in range(len(num_nbrs)):
        d[f'seed_nids{h}'] = batch.seed_nids[h].numpy()
        d[f'seed_times{h}'] = batch.seed_times[h].numpy()
        d[f'nbr_nids{h}'] = batch.nbr_nids[h].numpy()
        d[f'nbr_edge_x{h}'] = batch.nbr_edge_x[h].numpy()
        d[f'nbr_edge_time{h}'] = batch.nbr_edge_time[h].numpy()
    np.savez_compressed(os.path.join(HERE, f'nn_tgat_{name}.npz'), **d)
    print('tgat', name, out.shape, float(out.abs().max()))


def save_attention(name, S, k, node_dim, edge_dim, time_dim, heads, t_max, zero_bias, seed):
    g = torch.Generator().manual_seed(seed)
    att = TemporalAttention(heads, node_dim, edge_dim, time_dim).eval()
    te = Time2Vec(time_dim)
    randomise(att, seed + 1)
    if not zero_bias:
        with torch.no_grad():
            te.w.bias.copy_(0.3 * torch.randn(time_dim, generator=g))
    node_x = torch.randn(S, node_dim, generator=g)
    nbr_feat = torch.randn(S, k, node_dim, generator=g)
    edge_feat = torch.randn(S, k, edge_dim, generator=g)
    seed_t = torch.randint(0, t_max, (S,), generator=g)
    nbr_t = (seed_t[:, None] - torch.randint(1, max(2, t_max // 10), (S, k), generator=g)).clamp(min=0)
    nbr_t = torch.sort(nbr_t, 1)[0]
    nbr_id = torch.randint(0, 9000, (S, k), generator=g).int()
    npad = torch.randint(0, k + 1, (S,), generator=g)
    npad[0], npad[1] = k, 0  # one seed without any neighbour, one full
    pad = torch.arange(k)[None, :] < npad[:, None]  # left padding, as the sampler produces
    nbr_id[pad], nbr_t[pad], edge_feat[pad] = -1, 0, 0.0
    with torch.no_grad():
        out = att(node_x, te(torch.zeros(S)), edge_feat, nbr_feat, te(seed_t[:, None] - nbr_t),
                  nbr_id != -1)
    np.savez_compressed(
        os.path.join(HERE, f'nn_attn_{name}.npz'), node_x=node_x.numpy(), nbr_feat=nbr_feat.numpy(),
        edge_feat=edge_feat.numpy(), seed_t=seed_t.numpy(), nbr_t=nbr_t.numpy(),
        nbr_id=nbr_id.numpy(), out=out.numpy(), n_heads=np.int64(heads),
        **state_np(att, 'p.'), **state_np(te, 'p.time_encoder.'))
    print('attn', name, out.shape, float(out.abs().max()))


def save_attention_grad(name, S, k, node_dim, edge_dim, time_dim, heads, t_max, seed):
    """Gradients of loss = sum(out * G) through the reference TemporalAttention + Time2Vec by
    torch autograd (dropout 0: the stochastic mask is not reproducible across implementations)."""
    g = torch.Generator().manual_seed(seed)
    att = TemporalAttention(heads, node_dim, edge_dim, time_dim, dropout=0.0).train()
    te = Time2Vec(time_dim)
    randomise(att, seed + 1)
    with torch.no_grad():
        te.w.bias.copy_(0.3 * torch.randn(time_dim, generator=g))
    node_x = torch.randn(S, node_dim, generator=g, requires_grad=True)
    nbr_feat = torch.randn(S, k, node_dim, generator=g, requires_grad=True)
    edge_feat = torch.randn(S, k, edge_dim, generator=g)
    seed_t = torch.randint(0, t_max, (S,), generator=g)
    nbr_t = torch.sort((seed_t[:, None] - torch.randint(1, max(2, t_max // 10), (S, k), generator=g)).clamp(min=0), 1)[0]
    nbr_id = torch.randint(0, 9000, (S, k), generator=g).int()
    npad = torch.randint(0, k + 1, (S,), generator=g)
    npad[0], npad[1] = k, 0
    pad = torch.arange(k)[None, :] < npad[:, None]
    nbr_id[pad], nbr_t[pad] = -1, 0
    edge_feat[pad] = 0.0
    edge_feat.requires_grad_(True)
    G = torch.randn(S, att.out_dim, generator=g)
    out = att(node_x, te(torch.zeros(S)), edge_feat, nbr_feat, te(seed_t[:, None] - nbr_t), nbr_id != -1)
    (out * G).sum().backward()
    grads = {'g.' + n: p.grad.numpy() for n, p in att.named_parameters()}
    grads.update({'g.time_encoder.' + n: p.grad.numpy() for n, p in te.named_parameters()})
    np.savez_compressed(
        os.path.join(HERE, f'nn_attngrad_{name}.npz'), node_x=node_x.detach().numpy(),
        nbr_feat=nbr_feat.detach().numpy(), edge_feat=edge_feat.detach().numpy(), seed_t=seed_t.numpy(),
        nbr_t=nbr_t.numpy(), nbr_id=nbr_id.numpy(), G=G.numpy(), out=out.detach().numpy(),
        d_node_x=node_x.grad.numpy(), d_nbr_feat=nbr_feat.grad.numpy(), d_edge_feat=edge_feat.grad.numpy(),
        n_heads=np.int64(heads), **state_np(att, 'p.'), **state_np(te, 'p.time_encoder.'), **grads)
    print('attn grad', name, float(node_x.grad.abs().max()))


def save_tgat_grad(name, N, E, T, D, bs, num_nbrs, which, node_dim, time_dim, embed, heads):
    batch = sampled_batch(N, E, T, D, bs, num_nbrs, which, seed=len(name))
    torch.manual_seed(7)
    model = TGAT(node_dim=node_dim, edge_dim=D, time_dim=time_dim, embed_dim=embed,
                 num_layers=len(num_nbrs), n_heads=heads, dropout=0.0).train()
    randomise(model, 11)
    node_x = torch.randn(N, node_dim, generator=torch.Generator().manual_seed(3))
    out = model(node_x, batch.seed_nids, batch.seed_times, batch.nbr_nids, batch.nbr_edge_x,
                batch.nbr_edge_time)
    G = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    (out * G).sum().backward()
    d = dict(node_x=node_x.numpy(), out=out.detach().numpy(), G=G.numpy(), n_heads=np.int64(heads),
             num_layers=np.int64(len(num_nbrs)), **state_np(model))
    d.update({'g.' + n: p.grad.numpy() for n, p in model.named_parameters()})
    for h in range(len(num_nbrs)):
        d[f'seed_nids{h}'] = batch.seed_nids[h].numpy()
        d[f'seed_times{h}'] = batch.seed_times[h].numpy()
        d[f'nbr_nids{h}'] = batch.nbr_nids[h].numpy()
        d[f'nbr_edge_x{h}'] = batch.nbr_edge_x[h].numpy()
        d[f'nbr_edge_time{h}'] = batch.nbr_edge_time[h].numpy()
    np.savez_compressed(os.path.join(HERE, f'nn_tgatgrad_{name}.npz'), **d)
    print('tgat grad', name, out.shape)


def main():
    save_attention_grad('small', 40, 5, 3, 8, 10, 2, 5000, 21)
    save_attention_grad('pad_heads3', 33, 7, 4, 6, 9, 3, 800, 22)
    save_attention_grad('wide', 24, 20, 16, 24, 32, 2, 2_000_000, 23)
    save_tgat_grad('two_layer', 60, 1200, 400, 8, 20, [5, 5], 40, 3, 10, 12, 2)
    # name, S, k, node_dim, edge_dim, time_dim, heads, t_max, zero_bias, seed
    save_attention('small', 40, 5, 3, 8, 10, 2, 5000, True, 1)
    save_attention('pad_heads3', 33, 7, 4, 6, 9, 3, 800, False, 2)       # out 13 -> padded to 15
    save_attention('wiki_l1', 24, 20, 1, 172, 100, 2, 2_678_373, True, 3)  # key 273 -> 2*102
    save_attention('wiki_l2', 24, 20, 172, 172, 100, 2, 2_678_373, True, 4)  # key 444 -> 2*272
    # name, N, E, T, D, bs, num_nbrs, batch index, node_dim, time_dim, embed, heads, zero_bias
    save_tgat('two_layer', 60, 1200, 400, 8, 20, [5, 5], 40, 3, 10, 12, 2, True)
    save_tgat('one_layer_bias', 60, 900, 300, 6, 25, [6], 20, 1, 8, 10, 3, False)
    save_tgat('early_batch', 80, 600, 200, 4, 30, [4, 4], 1, 2, 6, 8, 2, True)  # mostly padded;
    # (hops must share k: tgat.py:139 reads the width of hop j-1 for every i)


if __name__ == '__main__':
    main()

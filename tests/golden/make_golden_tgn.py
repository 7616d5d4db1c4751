"""Golden fixtures for the TGN node-memory state machine, from the UNMODIFIED reference
TGNMemory (tgm/nn/encoder/tgn.py:80-251) with LastAggregator + IdentityMessage, driven the way
examples/linkproppred/tgn.py:60-124 drives it:  python tests/golden/make_golden_tgn.py

torch_geometric is not installed here; the import shim supplies `zeros` (in-place fill) and a
`scatter(reduce=max|mean)` restatement over torch.scatter_reduce_ (tests/golden/_ref_shim.py).
Writes tests/golden/tgn_*.npz (states) and tgngrad_*.npz (gradients) for the LastAggregator,
tgnmean_*.npz / tgnmeangrad_*.npz for the MeanAggregator."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shim import import_reference  # noqa: E402

import_reference()
from tgm.nn.encoder.tgn import IdentityMessage, LastAggregator, MeanAggregator, TGNMemory  # noqa: E402


def run(name, N, E, T, D, M, TD, bs, eval_from, seed, bias, mean=False):
    # Parity domain: within one batch no node may have two events with the same timestamp in the
    # same role.  TGNMemory._update_msg_store orders a node's events with `src.sort()`
    # (tgn.py:226), which is NOT stable on CPU, so LastAggregator's first-of-the-ties choice
    # (tgn.py:52) is implementation-defined there.  T > 0: unique timestamps.  T == 0: every
    # timestamp is shared by two consecutive edges with disjoint endpoints (ties across nodes).
    rng = np.random.default_rng(seed)
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    if T > 0:
        t = np.sort(rng.choice(T, E, replace=False))
    else:
        t = np.arange(E) // 2 * 3
        for i in range(1, E, 2):
            while len({src[i], dst[i], src[i - 1], dst[i - 1]}) < 4:
                src[i], dst[i] = rng.integers(0, N, 2)
                if dst[i - 1] == src[i - 1]:
                    dst[i - 1] = (src[i - 1] + 1) % N
    x = rng.standard_normal((E, D)).astype(np.float32)
    neg = rng.integers(0, N, E)
    torch.manual_seed(seed)
    mem = TGNMemory(N, D, M, TD, message_module=IdentityMessage(D, M, TD),
                    aggregator_module=MeanAggregator() if mean else LastAggregator())
    with torch.no_grad():
        for prm in mem.memory_updater.parameters():
            prm.copy_(torch.randn(prm.shape) * 0.3)
        if bias:
            mem.time_enc.w.bias.copy_(torch.randn(TD) * 0.3)
    mem.train()
    mem.reset_state()
    out = {}
    with torch.no_grad():
        for b, lo in enumerate(range(0, E, bs)):
            hi = min(lo + bs, E)
            if b == eval_from:
                mem.eval()  # flushes the message store into memory (tgn.py:245-251)
                out['flush_memory'] = mem.memory.numpy().copy()
                out['flush_last_update'] = mem.last_update.numpy().copy()
            s_, d_, t_ = (torch.from_numpy(a[lo:hi]).long() for a in (src, dst, t))
            n_id = torch.cat([s_, d_, torch.from_numpy(neg[lo:hi]).long()]).unique()
            z, lu = mem(n_id)
            out[f'b{b}_nid'] = n_id.numpy()
            out[f'b{b}_z'] = z.numpy().copy()
            out[f'b{b}_lu'] = lu.numpy().copy()
            mem.update_state(s_, d_, t_, torch.from_numpy(x[lo:hi]))
    out['final_memory'] = mem.memory.numpy().copy()
    out['final_last_update'] = mem.last_update.numpy().copy()
    sd = {'p.' + k: v.numpy() for k, v in mem.state_dict().items()
          if k not in ('memory', 'last_update', '_assoc')}
    np.savez_compressed(os.path.join(HERE, f"tgn{'mean' if mean else ''}_{name}.npz"), src=src.astype(np.int32),
                        dst=dst.astype(np.int32), t=t.astype(np.int64), x=x,
                        neg=neg.astype(np.int32), N=np.int64(N), bs=np.int64(bs),
                        eval_from=np.int64(eval_from), mean=np.int64(mean), **sd, **out)
    print(name, 'ok', float(np.abs(out['final_memory']).max()))


def run_grad(name, N, E, T, D, M, TD, bs, seed, bias, record_from, mean=False):
    """Training-mode gradients: the loop of examples/linkproppred/tgn.py:70-121 with the loss
    replaced by sum(z * G) for a recorded random G (forward, update_state, backward, detach); the
    .grad of every TGNMemory parameter is saved for the batches >= record_from (earlier batches
    only warm the memory and the message stores).  Writes tests/golden/tgngrad_*.npz."""
    rng = np.random.default_rng(seed)
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    t = np.sort(rng.choice(T, E, replace=False))
    x = rng.standard_normal((E, D)).astype(np.float32)
    neg = rng.integers(0, N, E)
    torch.manual_seed(seed)
    mem = TGNMemory(N, D, M, TD, message_module=IdentityMessage(D, M, TD),
                    aggregator_module=MeanAggregator() if mean else LastAggregator())
    with torch.no_grad():
        for prm in mem.memory_updater.parameters():
            prm.copy_(torch.randn(prm.shape) * 0.3)
        if bias:
            mem.time_enc.w.bias.copy_(torch.randn(TD) * 0.3)
    mem.train()
    mem.reset_state()
    out = {}
    for b, lo in enumerate(range(0, E, bs)):
        hi = min(lo + bs, E)
        s_, d_, t_ = (torch.from_numpy(a[lo:hi]).long() for a in (src, dst, t))
        n_id = torch.cat([s_, d_, torch.from_numpy(neg[lo:hi]).long()]).unique()
        mem.zero_grad()
        z, lu = mem(n_id)
        G = torch.from_numpy(rng.standard_normal(tuple(z.shape)).astype(np.float32))
        loss = (z * G).sum()
        mem.update_state(s_, d_, t_, torch.from_numpy(x[lo:hi]))
        loss.backward()
        mem.detach()
        if b >= record_from:
            out[f'b{b}_nid'] = n_id.numpy()
            out[f'b{b}_z'] = z.detach().numpy().copy()
            out[f'b{b}_G'] = G.numpy()
            for k, prm in mem.named_parameters():
                out[f'b{b}_g.{k}'] = prm.grad.numpy().copy()
    sd = {'p.' + k: v.numpy() for k, v in mem.state_dict().items()
          if k not in ('memory', 'last_update', '_assoc')}
    np.savez_compressed(os.path.join(HERE, f"tgn{'mean' if mean else ''}grad_{name}.npz"), src=src.astype(np.int32),
                        dst=dst.astype(np.int32), t=t.astype(np.int64), x=x,
                        neg=neg.astype(np.int32), N=np.int64(N), bs=np.int64(bs),
                        record_from=np.int64(record_from), mean=np.int64(mean), **sd, **out)
    print('grad', name, 'ok', max(float(np.abs(v).max()) for k, v in out.items() if '_g.' in k))


def main():
    # name, N, E, T, D, M, time_dim, bs, first eval-mode batch (-1: never), seed, t2v bias != 0
    run('train_small', 30, 400, 3000, 4, 8, 6, 20, -1, 1, False)
    run('train_ties', 12, 300, 0, 3, 6, 4, 25, -1, 2, True)         # timestamp ties across nodes
    run('train_then_eval', 40, 600, 5000, 5, 10, 8, 30, 12, 3, False)
    run('wiki_dims', 200, 600, 100000, 172, 100, 100, 200, 2, 4, False)
    # name, N, E, T, D, M, time_dim, bs, seed, t2v bias != 0, first recorded batch
    run_grad('small', 30, 300, 3000, 4, 8, 6, 20, 5, False, 3)
    run_grad('bias', 25, 300, 900, 3, 6, 4, 25, 6, True, 4)
    run_grad('c4_dims', 300, 1200, 2_000_000, 16, 100, 100, 200, 7, False, 3)
    # MeanAggregator (tgn.py:59-63): small node sets so that nodes collect several messages per batch
    run('small', 12, 400, 3000, 4, 8, 6, 20, -1, 8, True, mean=True)
    run('then_eval', 25, 600, 5000, 5, 10, 8, 30, 12, 9, False, mean=True)
    run_grad('small', 12, 300, 3000, 4, 8, 6, 20, 10, True, 3, mean=True)
    run_grad('c4_dims', 60, 1200, 2_000_000, 16, 100, 100, 200, 11, False, 3, mean=True)


if __name__ == '__main__':
    main()

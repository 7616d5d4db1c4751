"""GPU parity of the TGN memory with the MeanAggregator (tgm_tgn_set_aggregator(TGM_TGN_AGGR_MEAN):
tgn_message_mean_kernel, the append-only event log, the sin-sum form of the Time2Vec gradient).

"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.tgn_oracle import TGNMemoryOracle, tgn_memory_backward
from tests._golden import GOLDEN_DIR

pytestmark = pytest.mark.gpu

from tgm_b200.nn import IdentityMessage, MeanAggregator, TGNMemory  # noqa: E402

DEV = 'cuda:0'
TOL = 1e-5


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _close(got, want, what, rtol=5e-4):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    scale = max(1.0, float(np.abs(want).max()))
    err = float(np.abs(got - want).max())
    assert got.shape == want.shape and err <= rtol * scale, f'{what}: max err {err} (scale {scale})'


def _memory(p, N, D, M, TD):
    mem = TGNMemory(N, D, M, TD, message_module=IdentityMessage(D, M, TD),
                    aggregator_module=MeanAggregator())
    mem.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    return mem.to(DEV).train()


def _dims(z, p):
    return (int(z['N']), z['x'].shape[1], p['memory_updater.weight_hh'].shape[1],
            p['time_enc.w.bias'].shape[0])


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'tgnmean_*.npz'))),
                         ids=lambda p: os.path.basename(p)[8:-4])
def test_tgn_mean_memory_matches_reference_fixture(path):
    """forward(n_id) / update_state per batch, train -> eval flush in the middle, as
    tests/golden/make_golden_tgn.py drove the reference TGNMemory(aggregator_module=MeanAggregator())."""
    z = np.load(path)
    p = {k[2:]: z[k] for k in z.files if k.startswith('p.')}
    mem = _memory(p, *_dims(z, p))
    mem.reset_state()
    bs, eval_from, E = int(z['bs']), int(z['eval_from']), len(z['src'])
    with torch.no_grad():
        for b, lo in enumerate(range(0, E, bs)):
            hi = min(lo + bs, E)
            if b == eval_from:
                mem.eval()
                assert np.abs(mem.memory.cpu().numpy() - z['flush_memory']).max() <= TOL
                assert np.array_equal(mem.last_update.cpu().numpy(), z['flush_last_update'])
            zz, lu = mem(T(z[f'b{b}_nid']))
            assert np.abs(zz.cpu().numpy() - z[f'b{b}_z']).max() <= TOL, b
            assert np.array_equal(lu.cpu().numpy(), z[f'b{b}_lu']), b
            mem.update_state(T(z['src'][lo:hi]), T(z['dst'][lo:hi]), T(z['t'][lo:hi]), T(z['x'][lo:hi]))
    assert np.abs(mem.memory.cpu().numpy() - z['final_memory']).max() <= TOL
    assert np.array_equal(mem.last_update.cpu().numpy(), z['final_last_update'])


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'tgnmeangrad_*.npz'))),
                         ids=lambda p: os.path.basename(p)[12:-4])
def test_tgn_mean_memory_gradients_match_reference_autograd(path):
    z = np.load(path)
    p = {k[2:]: z[k] for k in z.files if k.startswith('p.')}
    mem = _memory(p, *_dims(z, p))
    mem.reset_state()
    bs, rec, E, checked = int(z['bs']), int(z['record_from']), len(z['src']), 0
    for b, lo in enumerate(range(0, E, bs)):
        hi = min(lo + bs, E)
        batch = [T(z[k][lo:hi]) for k in ('src', 'dst', 't', 'x')]
        if b < rec:
            mem.update_state(*batch)
            continue
        mem.zero_grad()
        zz, lu = mem(T(z[f'b{b}_nid']))
        assert zz.requires_grad and np.abs(zz.detach().cpu().numpy() - z[f'b{b}_z']).max() <= TOL, b
        loss = (zz * T(z[f'b{b}_G'])).sum()
        mem.update_state(*batch)  # the state moves on before backward, as in the reference loop
        loss.backward()
        for name, prm in mem.named_parameters():
            _close(prm.grad, z[f'b{b}_g.{name}'], f'batch {b} {name}')
        checked += 1
    assert checked >= 3


def test_tgn_mean_event_log_grows_and_restarts_after_a_flush():
    """More events than the log's initial capacity (65536) between two flushes: the log grows by
    doubling and every node still finds its last batch; a flush (train -> eval) starts it over."""
    rng = np.random.default_rng(31)
    N, D, M, TD, bs, nb = 400, 4, 8, 6, 4000, 20  # 80 000 events
    E = bs * nb
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    t = np.sort(rng.choice(5_000_000, E, replace=False))
    x = rng.standard_normal((E, D)).astype(np.float32)
    torch.manual_seed(5)
    mem = TGNMemory(N, D, M, TD, aggregator_module=MeanAggregator()).to(DEV).train()
    p = {k: v.detach().cpu().numpy().copy() for k, v in mem.state_dict().items()}
    orc = TGNMemoryOracle(N, D, M, TD, p, aggregator='mean')
    mem.reset_state()
    probe = np.arange(0, N, 7)
    with torch.no_grad():
        for b in range(nb):
            sl = slice(b * bs, (b + 1) * bs)
            mem.update_state(T(src[sl]), T(dst[sl]), T(t[sl]), T(x[sl]))
            orc.update_state(src[sl], dst[sl], t[sl], x[sl])
            if b in (0, nb // 2, nb - 1):
                zz, lu = mem(T(probe))
                wz, wlu = orc.forward(probe)
                assert np.abs(zz.cpu().numpy() - wz).max() <= TOL, b
                assert np.array_equal(lu.cpu().numpy(), wlu), b
        mem.eval()
        orc.train(False)
        assert np.abs(mem.memory.cpu().numpy() - orc.memory).max() <= TOL
        mem.train()
        orc.training = True
        sl = slice(0, bs)
        mem.update_state(T(src[sl]), T(dst[sl]), T(t[sl] + 6_000_000), T(x[sl]))
        orc.update_state(src[sl], dst[sl], t[sl] + 6_000_000, x[sl])
        zz, lu = mem(T(probe))
        wz, wlu = orc.forward(probe)
        assert np.abs(zz.cpu().numpy() - wz).max() <= TOL and np.array_equal(lu.cpu().numpy(), wlu)


def test_tgn_mean_gradients_vs_oracle_at_c4_dims_with_many_messages_per_node():
    rng = np.random.default_rng(17)
    N, E, D, M, TD, bs = 40, 800, 16, 100, 100, 200  # ~10 messages per node and role per batch
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    t = np.sort(rng.choice(2_000_000, E, replace=False))
    x = rng.standard_normal((E, D)).astype(np.float32)
    torch.manual_seed(9)
    mem = TGNMemory(N, D, M, TD, aggregator_module=MeanAggregator()).to(DEV).train()
    p = {k: v.detach().cpu().numpy().copy() for k, v in mem.state_dict().items()}
    orc = TGNMemoryOracle(N, D, M, TD, p, aggregator='mean')
    mem.reset_state()
    for lo in range(0, E, bs):
        sl = slice(lo, lo + bs)
        mem.update_state(T(src[sl]), T(dst[sl]), T(t[sl]), T(x[sl]))
        orc.update_state(src[sl], dst[sl], t[sl], x[sl])
    n_id = np.arange(N)
    G = rng.standard_normal((N, M)).astype(np.float32)
    zz, _ = mem(T(n_id))
    wz, _ = orc.forward(n_id)
    assert np.abs(zz.detach().cpu().numpy() - wz).max() <= TOL
    (zz * T(G)).sum().backward()
    want = tgn_memory_backward(orc, n_id, G)
    for name, prm in mem.named_parameters():
        _close(prm.grad, want[name], name)

"""GPU parity of the TGN training path: tgm_tgn_forward_saved / tgm_tgn_backward /
tgm_tgn_set_params and tgm_gae_backward / tgm_gae_set_params behind the autograd integration of
tgm_b200.nn.TGNMemory and GraphAttentionEmbedding.

The float64 oracles are pinned on the reference's autograd on CPU (tests/test_oracle_golden.py); the
same test bodies also run on CPU against an oracle-backed stand-in library
(tests/test_tgn_train_host_logic.py).  First hardware run: profiles/r1_tgn_train_gpu_tests.log.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.tgn_oracle import (TGNMemoryOracle, graph_attention_embedding,
                               graph_attention_embedding_backward, tgn_memory_backward)
from tests._golden import GOLDEN_DIR

pytestmark = pytest.mark.gpu

from tgm_b200.nn import GraphAttentionEmbedding, TGNMemory, Time2Vec  # noqa: E402

DEV = 'cuda:0'
TOL = 1e-5


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _close(got, want, what, rtol=5e-4):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    scale = max(1.0, float(np.abs(want).max()))
    err = float(np.abs(got - want).max())
    assert got.shape == want.shape and err <= rtol * scale, f'{what}: max err {err} (scale {scale})'


def _memory_from(p, N, D, M, TD):
    mem = TGNMemory(N, D, M, TD)
    mem.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    return mem.to(DEV).train()


# ---- memory updater: gradients of the reference's autograd -----------------------------------------
@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'tgngrad_*.npz'))),
                         ids=lambda p: os.path.basename(p)[8:-4])
def test_tgn_memory_gradients_match_reference_autograd(path):
    """The loop of tests/golden/make_golden_tgn.py::run_grad (= examples/linkproppred/tgn.py:70-121
    with loss = sum(z * G)): forward, update_state, THEN backward; every TGNMemory parameter's .grad
    against the unmodified reference's."""
    z = np.load(path)
    p = {k[2:]: z[k] for k in z.files if k.startswith('p.')}
    N, bs, rec = int(z['N']), int(z['bs']), int(z['record_from'])
    D, M, TD = z['x'].shape[1], p['memory_updater.weight_hh'].shape[1], p['time_enc.w.bias'].shape[0]
    mem = _memory_from(p, N, D, M, TD)
    mem.reset_state()
    E, checked = len(z['src']), 0
    for b, lo in enumerate(range(0, E, bs)):
        hi = min(lo + bs, E)
        batch = [T(z[k][lo:hi]) for k in ('src', 'dst', 't', 'x')]
        if b < rec:
            with torch.no_grad():
                mem.update_state(*batch)
            continue
        mem.zero_grad()
        zz, lu = mem(T(z[f'b{b}_nid']))
        assert zz.requires_grad and not lu.requires_grad
        assert np.abs(zz.detach().cpu().numpy() - z[f'b{b}_z']).max() <= TOL, b
        loss = (zz * T(z[f'b{b}_G'])).sum()
        mem.update_state(*batch)   # the state moves on before backward, as in the reference loop
        loss.backward()
        mem.detach()
        for name, prm in mem.named_parameters():
            _close(prm.grad, z[f'b{b}_g.{name}'], f'batch {b} {name}')
        checked += 1
    assert checked >= 3


def test_tgn_memory_parameter_refresh_keeps_state_and_message_stores():
    """An optimizer step changes the parameters between batches: the handle refreshes its copies
    in place (tgm_tgn_set_params); memory, last_update and the stored messages carry over."""
    rng = np.random.default_rng(21)
    N, E, D, M, TD, bs = 60, 400, 5, 12, 8, 40
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    t = np.sort(rng.choice(50_000, E, replace=False))
    x = rng.standard_normal((E, D)).astype(np.float32)
    torch.manual_seed(2)
    mem = TGNMemory(N, D, M, TD).to(DEV).train()
    p = {k: v.detach().cpu().numpy().copy() for k, v in mem.state_dict().items()}
    orc = TGNMemoryOracle(N, D, M, TD, p)
    mem.reset_state()
    for b, lo in enumerate(range(0, E, bs)):
        sl = slice(lo, lo + bs)
        n_id = np.unique(np.concatenate([src[sl], dst[sl]]))
        with torch.no_grad():
            zz, lu = mem(T(n_id))
        wz, wlu = orc.forward(n_id)
        assert np.abs(zz.cpu().numpy() - wz).max() <= TOL and np.array_equal(lu.cpu().numpy(), wlu), b
        with torch.no_grad():
            mem.update_state(T(src[sl]), T(dst[sl]), T(t[sl]), T(x[sl]))
        orc.update_state(src[sl], dst[sl], t[sl], x[sl])
        if b % 3 == 2:  # "optimizer step"
            with torch.no_grad():
                for prm in mem.parameters():
                    prm.add_(torch.randn_like(prm) * 0.02)
            orc.p = {k: v.detach().cpu().numpy().copy() for k, v in mem.state_dict().items()}
    assert np.abs(mem.memory.cpu().numpy() - orc.memory).max() <= TOL
    assert np.array_equal(mem.last_update.cpu().numpy(), orc.last_update)


def test_tgn_memory_gradients_vs_oracle_with_nodes_without_messages():
    """n_id mixes nodes with a source-store message, a destination-store message, both, and none
    (their GRU input is the zero row and Time2Vec gets no gradient from them)."""
    rng = np.random.default_rng(5)
    N, E, D, M, TD, bs = 200, 300, 16, 100, 100, 100
    src, dst = rng.integers(0, N // 2, E), rng.integers(0, N // 2, E)   # upper half: never touched
    t = np.sort(rng.choice(2_000_000, E, replace=False))
    x = rng.standard_normal((E, D)).astype(np.float32)
    torch.manual_seed(4)
    mem = TGNMemory(N, D, M, TD).to(DEV).train()
    p = {k: v.detach().cpu().numpy().copy() for k, v in mem.state_dict().items()}
    orc = TGNMemoryOracle(N, D, M, TD, p)
    mem.reset_state()
    for lo in range(0, E, bs):
        sl = slice(lo, lo + bs)
        with torch.no_grad():
            mem.update_state(T(src[sl]), T(dst[sl]), T(t[sl]), T(x[sl]))
        orc.update_state(src[sl], dst[sl], t[sl], x[sl])
    n_id = np.arange(0, N, 3)
    G = rng.standard_normal((len(n_id), M)).astype(np.float32)
    zz, _ = mem(T(n_id))
    (zz * T(G)).sum().backward()
    want = tgn_memory_backward(orc, n_id, G)
    for name, prm in mem.named_parameters():
        _close(prm.grad, want[name], name)


# ---- embedding: gradients vs the float64 oracle (parity UNPINNED: third-party convolution) -----------
def _gae_case(rng, n, m, M, Z, D, TD, bias=True):
    torch.manual_seed(3)
    te = Time2Vec(TD)
    enc = GraphAttentionEmbedding(in_channels=M, out_channels=Z, msg_dim=D, time_enc=te)
    enc.conv.dropout = 0.0
    if bias:
        with torch.no_grad():
            te.w.bias.copy_(torch.randn(TD) * 0.3)
    enc = enc.to(DEV).train()
    p = {k: v.detach().cpu().numpy() for k, v in enc.state_dict().items()}
    x = rng.standard_normal((n, M)).astype(np.float32)
    lu = rng.integers(0, 2_000_000, n)
    src = rng.integers(0, n, m)
    dst = rng.integers(0, max(1, n // 2), m)
    if m > 10:
        dst[: m // 4] = 3  # a hub: one target with a quarter of all edges
    t = rng.integers(0, 2_000_000, m)
    msg = rng.standard_normal((m, D)).astype(np.float32)
    return enc, p, x, lu, np.stack([src, dst]), t, msg


@pytest.mark.parametrize('dims', [(700, 6000, 100, 100, 172, 100), (40, 300, 5, 100, 7, 2),
                                  (33, 65, 8, 6, 0, 3), (50, 0, 16, 8, 4, 4)],
                         ids=['tgn_example_wiki', 'reference_test_dims', 'no_msg_feats', 'no_edges'])
def test_graph_attention_embedding_gradients_vs_oracle(dims):
    n, m, M, Z, D, TD = dims
    rng = np.random.default_rng(n + 1)
    enc, p, x, lu, ei, t, msg = _gae_case(rng, n, m, M, Z, D, TD)
    xt = T(x).requires_grad_()
    out = enc(xt, T(lu), T(ei), T(t), T(msg))
    assert out.requires_grad
    assert np.abs(out.detach().cpu().numpy() - graph_attention_embedding(p, 2, x, lu, ei, t, msg)).max() <= TOL
    G = rng.standard_normal((n, Z)).astype(np.float32)
    (out * T(G)).sum().backward()
    want = graph_attention_embedding_backward(p, 2, x, lu, ei, t, msg, G)
    _close(xt.grad, want['x'], 'x')
    for name, prm in enc.named_parameters():
        _close(prm.grad, want[name], name)


def test_tgn_training_step_memory_into_embedding():
    """One training step of examples/linkproppred/tgn.py:100-118 without the decoder: z =
    embedding(memory(n_id)); the embedding's d x is the memory's d_memory, and the shared
    Time2Vec accumulates the gradients of both modules.  Then an Adam step refreshes both
    handles in place and the next forward uses the new parameters."""
    rng = np.random.default_rng(8)
    N, E, D, M, TD, Z, bs = 150, 400, 16, 100, 100, 100, 100
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    t = np.sort(rng.choice(2_000_000, E, replace=False))
    x = rng.standard_normal((E, D)).astype(np.float32)
    torch.manual_seed(6)
    mem = TGNMemory(N, D, M, TD).to(DEV).train()
    enc = GraphAttentionEmbedding(M, Z, D, mem.time_enc)
    enc.conv.dropout = 0.0
    enc = enc.to(DEV).train()
    params = {id(q): q for q in list(mem.parameters()) + list(enc.parameters())}
    opt = torch.optim.Adam(params.values(), lr=1e-3)
    snap = lambda mod: {k: v.detach().cpu().numpy().copy() for k, v in mod.state_dict().items()}
    orc = TGNMemoryOracle(N, D, M, TD, snap(mem))
    mem.reset_state()
    for lo in range(0, 300, bs):
        sl = slice(lo, lo + bs)
        with torch.no_grad():
            mem.update_state(T(src[sl]), T(dst[sl]), T(t[sl]), T(x[sl]))
        orc.update_state(src[sl], dst[sl], t[sl], x[sl])
    sl = slice(300, 400)
    n_id = np.unique(np.concatenate([src[sl], dst[sl]]))
    n = len(n_id)
    m = 5 * n
    ei = np.stack([rng.integers(0, n, m), rng.integers(0, n, m)])
    et = rng.integers(0, 2_000_000, m)
    emsg = rng.standard_normal((m, D)).astype(np.float32)
    G = rng.standard_normal((n, Z)).astype(np.float32)

    opt.zero_grad()
    zm, lu = mem(T(n_id))
    zz = enc(zm, lu, T(ei), T(et), T(emsg))
    loss = (zz * T(G)).sum()
    mem.update_state(T(src[sl]), T(dst[sl]), T(t[sl]), T(x[sl]))
    loss.backward()

    pe = snap(enc)
    wz, wlu = orc.forward(n_id)
    ge = graph_attention_embedding_backward(pe, 2, wz, wlu, ei, et, emsg, G)
    gm = tgn_memory_backward(orc, n_id, ge['x'])
    for name, prm in mem.named_parameters():
        want = gm[name] + (ge[name] if name.startswith('time_enc') else 0)
        _close(prm.grad, want, 'memory ' + name)
    for name, prm in enc.named_parameters():
        if not name.startswith('time_enc'):
            _close(prm.grad, ge[name], 'embedding ' + name)

    opt.step()
    orc.update_state(src[sl], dst[sl], t[sl], x[sl])
    orc.p = snap(mem)
    with torch.no_grad():
        zm2, lu2 = mem(T(n_id))
        zz2 = enc.eval()(zm2, lu2, T(ei), T(et), T(emsg))
    wz2, wlu2 = orc.forward(n_id)
    assert np.abs(zm2.cpu().numpy() - wz2).max() <= TOL and np.array_equal(lu2.cpu().numpy(), wlu2)
    want2 = graph_attention_embedding(snap(enc), 2, wz2, wlu2, ei, et, emsg)
    assert np.abs(zz2.cpu().numpy() - want2).max() <= 2 * TOL

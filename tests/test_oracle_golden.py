"""Pin the CPU oracle: known answers of the reference's own unit tests + reference-generated
fixtures (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle.c_oracle import CRing
from oracle.recency_oracle import PADDED_NODE_ID, RingOracle, masked_mean, stateless_sample
from tests._golden import Golden, assert_hop_equal, golden_files, golden_ids


@pytest.mark.parametrize('impl', [RingOracle, CRing], ids=['numpy', 'c'])
@pytest.mark.parametrize('path', golden_files(), ids=golden_ids())
def test_ring_oracle_matches_reference_fixture(path, impl):
    g = Golden(path)
    ring = impl(g.N, g.num_nbrs, g.D, g.directed)
    for ep in range(g.epochs):
        if ep:
            ring.reset_state()
        for b, lo, hi in g.batches():
            s, q = g.seeds(lo, hi)
            hops = ring.hook_call(s, q, g.src[lo:hi], g.dst[lo:hi], g.t[lo:hi],
                                  None if g.x is None else g.x[lo:hi])
            for h, got in enumerate(hops):
                assert_hop_equal(got, g.expect(ep, b, h), f'ep{ep} batch{b} hop{h}')
    assert np.array_equal(ring.ids, g.z['final_ids'])
    assert np.array_equal(ring.times, g.z['final_times'])
    assert np.array_equal(ring.feats, g.z['final_feats'])
    assert np.array_equal(ring.write_pos, g.z['final_write_pos'])


@pytest.mark.parametrize('path', golden_files(), ids=golden_ids())
def test_stateless_oracle_matches_reference_fixture(path):
    g = Golden(path)
    res = stateless_sample(g.src, g.dst, g.t, g.x, g.bs, g.num_nbrs,
                           lambda b, lo, hi: g.seeds(lo, hi), g.directed)
    for b, hops in enumerate(res):
        for h, got in enumerate(hops):
            assert_hop_equal(got, g.expect(0, b, h), f'batch{b} hop{h}')


def _run(src, dst, t, x, N, bs, num_nbrs, directed=False):
    ring = RingOracle(N, num_nbrs, x.shape[1], directed)
    out = []
    for lo in range(0, len(src), bs):
        hi = lo + bs
        s = np.concatenate([src[lo:hi], dst[lo:hi]]).astype(np.int32)
        q = np.concatenate([t[lo:hi], t[lo:hi]]).astype(np.int64)
        out.append(ring.hook_call(s, q, src[lo:hi], dst[lo:hi], t[lo:hi], x[lo:hi]))
    return out


def test_kat_alice_bob_1hop():
    """Values asserted by test/unit/test_hooks/test_recency_nbr_hook.py:344-417."""
    src, dst = np.array([0, 0, 2, 2]), np.array([1, 2, 3, 0])
    t, x = np.array([1, 2, 3, 4]), np.array([[1], [2], [5], [2]], np.float32)
    b = _run(src, dst, t, x, 4, 1, [1])
    nid = [h[0][2][:, 0].tolist() for h in b]
    nt = [h[0][3][:, 0].tolist() for h in b]
    nx = [h[0][4][:, 0, 0].tolist() for h in b]
    assert nid == [[-1, -1], [1, -1], [0, -1], [3, 2]]
    assert nt == [[0, 0], [1, 0], [2, 0], [3, 2]]
    assert nx == [[0.0, 0.0], [1.0, 0.0], [2.0, 0.0], [5.0, 2.0]]


def test_kat_alice_bob_1hop_directed():
    """test_recency_nbr_hook.py:420-494: batch 3 sees no neighbour for node 2 when directed."""
    src, dst = np.array([0, 0, 2, 2]), np.array([1, 2, 3, 0])
    t, x = np.array([1, 2, 3, 4]), np.array([[1], [2], [5], [2]], np.float32)
    b = _run(src, dst, t, x, 4, 1, [1], directed=True)
    assert [h[0][2][:, 0].tolist() for h in b] == [[-1, -1], [1, -1], [-1, -1], [3, 2]]
    assert [h[0][3][:, 0].tolist() for h in b] == [[0, 0], [1, 0], [0, 0], [3, 2]]


def test_kat_star_exceeds_buffer():
    """test_recency_nbr_hook.py:521-566: node 0 always reports its two most recent spokes."""
    src, dst = np.zeros(100, int), np.arange(1, 101)
    t, x = np.arange(100), np.arange(1, 101, dtype=np.float32)[:, None]
    b = _run(src, dst, t, x, 101, 2, [2])
    assert b[0][0][2][0].tolist() == [PADDED_NODE_ID, PADDED_NODE_ID]
    assert b[1][0][2][0].tolist() == [1, 2] and b[1][0][3][0].tolist() == [0, 1]
    for hops in b[2:]:
        nid, nt, nx = hops[0][2][0], hops[0][3][0], hops[0][4][0, :, 0]
        assert (nid == nt + 1).all() and (nx == nt + 1).all()


def test_kat_two_hop_pushed_out_of_cache():
    """test_recency_nbr_hook.py:593-678 (the comments there give the expected ids)."""
    src, dst = np.array([0, 1, 3, 4, 5, 5]), np.array([1, 2, 2, 2, 0, 2])
    t, x = np.arange(1, 7), np.array([[1], [3], [5], [6], [5], [7]], np.float32)
    b = _run(src, dst, t, x, 6, 1, [1, 1])
    hop0 = [h[0][2][:, 0].tolist() for h in b]
    hop1 = [h[1][2][:, 0].tolist() for h in b]
    assert hop0 == [[-1, -1], [0, -1], [-1, 1], [-1, 3], [-1, 1], [0, 4]]
    assert hop1 == [[-1, -1], [-1, -1], [-1, -1], [-1, -1], [-1, -1], [-1, -1]]


def test_constructor_errors():
    """recency.py:66-69."""
    with pytest.raises(ValueError):
        RingOracle(4, [], 0)
    with pytest.raises(ValueError):
        RingOracle(4, [0], 0)


def test_masked_mean_kat():
    z = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    ids = np.array([[-1, 4, 5], [-1, -1, -1]], np.int32)
    out = masked_mean(z, ids)
    assert np.allclose(out[0], (z[0, 1] + z[0, 2]) / 2) and (out[1] == 0).all()


@pytest.mark.parametrize('path', golden_files(), ids=golden_ids())
def test_c_run_stream_checksums_match_fixture(path):
    """ring_run_stream (the full-size checker) reproduces the fixture's outputs and its running
    checksums equal checksum_np over the concatenated per-batch expectations."""
    g = Golden(path)
    if g.neg is not None:
        pytest.skip('run_stream drives the [src | dst] seed layout only')
    ring = CRing(g.N, g.num_nbrs, g.D, g.directed)
    slots, csum, outs = ring.run_stream(g.src, g.dst, g.t, g.x, 0, g.E, g.bs, keep_hop0=True)
    nb = sum(1 for _ in g.batches())
    want_slots = 0
    for h in range(len(g.num_nbrs)):
        cat = [np.concatenate([g.expect(0, b, h)[i] for b in range(nb)]) for i in (2, 3, 4)]
        want_slots += cat[0].size
        assert int(csum[h, 0]) == c_oracle.checksum_np(cat[0])
        assert int(csum[h, 1]) == c_oracle.checksum_np(cat[1])
        if g.D:
            assert int(csum[h, 2]) == c_oracle.checksum_np(cat[2])
        if h == 0:
            for got, want in zip(outs, cat):
                assert np.array_equal(got, want)
    assert slots == want_slots


def test_c_masked_mean_equals_numpy_oracle():
    rng = np.random.default_rng(3)
    z = rng.standard_normal((50, 7, 12)).astype(np.float32)
    ids = np.where(rng.random((50, 7)) < 0.4, -1, rng.integers(0, 9, (50, 7))).astype(np.int32)
    ids[0] = -1
    assert np.array_equal(c_oracle.masked_mean(z, ids), masked_mean(z, ids))


def test_c_oracle_equals_numpy_oracle_on_a_larger_stream():
    rng = np.random.default_rng(11)
    N, E, T, D, bs, nn = 500, 20000, 400, 4, 200, [20, 5]
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    src = np.where(rng.random(E) < 0.2, 3, src)  # hot node: > B pushes per batch
    t = np.sort(rng.integers(0, T, E))
    x = rng.standard_normal((E, D)).astype(np.float32)
    a, b = RingOracle(N, nn, D), CRing(N, nn, D)
    for lo in range(0, E, bs):
        hi = lo + bs
        s = np.concatenate([src[lo:hi], dst[lo:hi]]).astype(np.int32)
        q = np.concatenate([t[lo:hi], t[lo:hi]]).astype(np.int64)
        ha = a.hook_call(s, q, src[lo:hi], dst[lo:hi], t[lo:hi], x[lo:hi])
        hb = b.hook_call(s, q, src[lo:hi], dst[lo:hi], t[lo:hi], x[lo:hi])
        for h, (u, v) in enumerate(zip(ha, hb)):
            assert_hop_equal(v, u, f'edge {lo} hop{h}')
    assert np.array_equal(a.ids, b.ids) and np.array_equal(a.write_pos, b.write_pos)


def test_time2vec_oracle_pins_torch_linear():
    """The Time2Vec oracle against the live torch module arithmetic (time_encoding.py:12-24).
    b = 0 (shipped init): torch's argument equals the oracle's bit for bit and the cosine agrees
    to 1e-6.  b != 0: every torch argument equals the fused or the two-rounding form (which one
    depends on the BLAS path), so only that pair is a well-posed target."""
    import torch
    from oracle.recency_oracle import time2vec
    d = 100
    w = (1.0 / 10 ** np.linspace(0, 9, d)).astype(np.float32)
    rng = np.random.default_rng(0)
    dt = rng.integers(0, 2_700_000, 2000).astype(np.int64)
    lin = torch.nn.Linear(1, d)
    for b in (np.zeros(d, np.float32), rng.standard_normal(d).astype(np.float32)):
        with torch.no_grad():
            lin.weight.copy_(torch.from_numpy(w).reshape(d, 1))
            lin.bias.copy_(torch.from_numpy(b))
            arg = lin(torch.from_numpy(dt).float().unsqueeze(-1))
            out = torch.cos(arg).numpy()
        arg = arg.numpy()
        x = dt.astype(np.float32)[:, None]
        fused = (x.astype(np.float64) * w.astype(np.float64) + b.astype(np.float64)).astype(np.float32)
        twice = (x * w).astype(np.float32) + b
        assert ((arg == fused) | (arg == twice)).all()
        if not b.any():
            assert (arg == fused).all() and (fused == twice).all()
            assert np.abs(out - time2vec(dt, w, b)).max() <= 1e-6


# ---- aggregation modules: numpy oracle vs fixtures from the live reference modules ---------------
import glob  # noqa: E402
import os  # noqa: E402

from oracle import nn_oracle  # noqa: E402
from tests._golden import GOLDEN_DIR  # noqa: E402


def _params(z):
    return {k[2:]: z[k] for k in z.files if k.startswith('p.')}


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'nn_attn_*.npz'))),
                         ids=lambda p: os.path.basename(p)[8:-4])
def test_attention_oracle_matches_reference_module(path):
    z = np.load(path)
    p = _params(z)
    S = z['node_x'].shape[0]
    tf0 = nn_oracle._t2v(p, 'time_encoder.', np.zeros(S, np.int64))
    tfn = nn_oracle._t2v(p, 'time_encoder.', z['seed_t'][:, None] - z['nbr_t'])
    out = nn_oracle.temporal_attention(p, '', int(z['n_heads']), z['node_x'], tf0, z['edge_feat'],
                                       z['nbr_feat'], tfn, z['nbr_id'] != -1)
    assert out.dtype == np.float32 and out.shape == z['out'].shape
    assert np.abs(out - z['out']).max() <= 5e-6


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'nn_tgat_*.npz'))),
                         ids=lambda p: os.path.basename(p)[8:-4])
def test_tgat_oracle_matches_reference_module(path):
    z = np.load(path)
    L = int(z['num_layers'])
    out = nn_oracle.tgat_forward(
        _params(z), L, int(z['n_heads']), z['node_x'],
        [z[f'seed_nids{h}'] for h in range(L)], [z[f'seed_times{h}'] for h in range(L)],
        [z[f'nbr_nids{h}'] for h in range(L)], [z[f'nbr_edge_x{h}'] for h in range(L)],
        [z[f'nbr_edge_time{h}'] for h in range(L)])
    assert np.abs(out - z['out']).max() <= 5e-6


# ---- uniform sampler oracle vs reference fixtures ------------------------------------------------
from oracle.recency_oracle import uniform_sample_deterministic  # noqa: E402


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'uniform_*.npz'))),
                         ids=lambda p: os.path.basename(p)[8:-4])
def test_uniform_oracle_matches_reference_fixture(path):
    """NeighborSamplerHook (uniform.py:87-142): every hop samples the history strictly before the
    batch's earliest time; fixtures have k >= every degree, so the reference is deterministic."""
    z = np.load(path)
    src, dst, t = z['src'], z['dst'], z['t']
    x = z['x'] if int(z['has_x']) else None
    bs, nn, directed = int(z['bs']), [int(v) for v in z['num_nbrs']], bool(int(z['directed']))
    for b, lo in enumerate(range(0, len(src), bs)):
        hi = min(lo + bs, len(src))
        e_hi = int(np.searchsorted(t, t[lo:hi].min() - 1, 'right'))  # end_time = min t - 1
        seeds = np.concatenate([src[lo:hi], dst[lo:hi]])
        for h, k in enumerate(nn):
            if h:
                seeds = nid.reshape(-1)
            nid, nt, nx, exact = uniform_sample_deterministic(src, dst, t, x, 0, e_hi, seeds, k,
                                                              directed)
            assert exact.all()
            assert np.array_equal(nid, z[f'b{b}_h{h}_nid'])
            assert np.array_equal(nt, z[f'b{b}_h{h}_nt'])
            assert np.array_equal(nx, z[f'b{b}_h{h}_nx'])
            assert np.array_equal(seeds.astype(np.int32), z[f'b{b}_h{h}_seed'])


# ---- TGN node memory oracle vs fixtures from the live reference TGNMemory ------------------------
from oracle.tgn_oracle import TGNMemoryOracle  # noqa: E402


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'tgn_*.npz'))),
                         ids=lambda p: os.path.basename(p)[4:-4])
def test_tgn_memory_oracle_matches_reference(path):
    z = np.load(path)
    p = _params(z)
    N, bs, eval_from = int(z['N']), int(z['bs']), int(z['eval_from'])
    D = z['x'].shape[1]
    M = p['memory_updater.weight_hh'].shape[1]
    TD = p['time_enc.w.bias'].shape[0]
    mem = TGNMemoryOracle(N, D, M, TD, p)
    E = len(z['src'])
    for b, lo in enumerate(range(0, E, bs)):
        hi = min(lo + bs, E)
        if b == eval_from:
            mem.train(False)
            assert np.abs(mem.memory - z['flush_memory']).max() <= 5e-6
            assert np.array_equal(mem.last_update, z['flush_last_update'])
        zz, lu = mem.forward(z[f'b{b}_nid'])
        assert np.abs(zz - z[f'b{b}_z']).max() <= 5e-6, b
        assert np.array_equal(lu, z[f'b{b}_lu']), b
        mem.update_state(z['src'][lo:hi], z['dst'][lo:hi], z['t'][lo:hi], z['x'][lo:hi])
    assert np.abs(mem.memory - z['final_memory']).max() <= 5e-6
    assert np.array_equal(mem.last_update, z['final_last_update'])


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'nn_dygformer_*.npz'))),
                         ids=lambda p: os.path.basename(p)[13:-4])
def test_dygformer_oracle_matches_reference_module(path):
    z = np.load(path)
    zs, zd = nn_oracle.dygformer_forward(
        _params(z), int(z['patch_size']), int(z['num_layers']), int(z['num_heads']), z['node_x'],
        np.stack([z['src'], z['dst']]), z['t'], z['nbrs'], z['nt'], z['ef'])
    assert np.abs(zs - z['z_src']).max() <= 5e-6 and np.abs(zd - z['z_dst']).max() <= 5e-6


# ---- GraphAttentionEmbedding / TransformerConv (parity UNPINNED: third-party arithmetic) -----------
def _torch_transformer_conv(p, heads, x, ei, ea):
    """Independent restatement in torch ops (scatter_reduce/index_add), written from
    torch_geometric 2.6.1's TransformerConv.forward/message and utils.softmax; two restatements
    agreeing is a consistency check of the oracle, not a pin against the real package."""
    import math

    import torch
    x, ea = torch.from_numpy(x), torch.from_numpy(ea)
    j, i = torch.from_numpy(ei[0]), torch.from_numpy(ei[1])
    W = lambda n: torch.from_numpy(p[f'conv.{n}'])
    n, HC = x.shape[0], p['conv.lin_query.weight'].shape[0]
    C = HC // heads
    q = torch.nn.functional.linear(x, W('lin_query.weight'), W('lin_query.bias')).view(n, heads, C)
    k = torch.nn.functional.linear(x, W('lin_key.weight'), W('lin_key.bias')).view(n, heads, C)
    v = torch.nn.functional.linear(x, W('lin_value.weight'), W('lin_value.bias')).view(n, heads, C)
    e = torch.nn.functional.linear(ea, W('lin_edge.weight')).view(-1, heads, C)
    alpha = (q[i] * (k[j] + e)).sum(-1) / math.sqrt(C)
    mx = torch.full((n, heads), float('-inf')).scatter_reduce(0, i[:, None].expand(-1, heads), alpha,
                                                             'amax', include_self=True)
    ex = (alpha - mx[i]).exp()
    den = torch.zeros(n, heads).index_add_(0, i, ex) + 1e-16
    a = ex / den[i]
    out = torch.zeros(n, heads, C).index_add_(0, i, (v[j] + e) * a[:, :, None])
    return (out.view(n, HC) + torch.nn.functional.linear(x, W('lin_skip.weight'), W('lin_skip.bias'))).numpy()


def test_transformer_conv_oracle_agrees_with_an_independent_torch_restatement():
    from oracle.tgn_oracle import graph_attention_embedding, transformer_conv
    rng = np.random.default_rng(5)
    n, m, IN, HC, H, D, TD = 60, 400, 12, 16, 2, 5, 6
    p = {f'conv.{nm}.weight': rng.standard_normal((HC, IN)).astype(np.float32) * 0.3
         for nm in ('lin_query', 'lin_key', 'lin_value', 'lin_skip')}
    p.update({f'conv.{nm}.bias': rng.standard_normal(HC).astype(np.float32) * 0.1
              for nm in ('lin_query', 'lin_key', 'lin_value', 'lin_skip')})
    p['conv.lin_edge.weight'] = rng.standard_normal((HC, TD + D)).astype(np.float32) * 0.3
    p['time_enc.w.weight'] = (1 / 10 ** np.linspace(0, 4, TD)).astype(np.float32).reshape(TD, 1)
    p['time_enc.w.bias'] = np.zeros(TD, np.float32)
    x = rng.standard_normal((n, IN)).astype(np.float32)
    ei = np.stack([rng.integers(0, n, m), rng.integers(0, n // 2, m)])  # half the nodes: no in-edges
    ea = rng.standard_normal((m, TD + D)).astype(np.float32)
    got = transformer_conv(p, 'conv.', H, x, ei, ea)
    want = _torch_transformer_conv(p, H, x, ei, ea)
    assert got.shape == (n, HC) and np.abs(got - want).max() <= 2e-6
    # nodes without incoming edges reduce to the skip projection
    skip = x @ p['conv.lin_skip.weight'].T + p['conv.lin_skip.bias']
    assert np.abs(got[n // 2:] - skip[n // 2:]).max() <= 1e-6
    # the embedding wrapper builds edge_attr = [Time2Vec(last_update[src] - t) | msg]
    lu, t = rng.integers(0, 1000, n), rng.integers(0, 1000, m)
    msg = rng.standard_normal((m, D)).astype(np.float32)
    z = graph_attention_embedding(p, H, x, lu, ei, t, msg)
    enc = np.cos(((lu[ei[0]] - t).astype(np.float32)[:, None] * p['time_enc.w.weight'].reshape(1, -1)))
    want = _torch_transformer_conv(p, H, x, ei, np.concatenate([enc.astype(np.float32), msg], 1))
    assert np.abs(z - want).max() <= 1e-5


# ---- DyGFormer gradients: the hand-derived backward vs the reference's autograd -----------------------
@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'nn_dyggrad_*.npz'))),
                         ids=lambda p: os.path.basename(p)[11:-4])
def test_dygformer_backward_oracle_matches_reference_autograd(path):
    """oracle/nn_oracle.py::dygformer_backward (float64 chain rule) against .grad of every parameter
    of the unmodified reference module after loss.backward() (tests/golden/make_golden_dygformer.py
    ::run_grad)."""
    from oracle import nn_oracle
    z = np.load(path)
    p = {k[2:]: z[k] for k in z.files if k.startswith('p.')}
    g = nn_oracle.dygformer_backward(p, int(z['patch_size']), int(z['num_layers']), int(z['num_heads']),
                                     z['node_x'], np.stack([z['src'], z['dst']]), z['t'], z['nbrs'],
                                     z['nt'], z['ef'], z['G_src'], z['G_dst'])
    names = [k[2:] for k in z.files if k.startswith('g.')]
    assert set(names) == set(g) and len(names) >= 20
    for name in names:
        want = z['g.' + name]
        assert g[name].shape == want.shape, name
        assert np.abs(g[name] - want).max() <= 5e-5 * max(1.0, np.abs(want).max()), name


# ---- TGN gradients: hand-derived backward of the memory updater and of the embedding ------------------
@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'tgngrad_*.npz'))),
                         ids=lambda p: os.path.basename(p)[8:-4])
def test_tgn_memory_backward_oracle_matches_reference_autograd(path):
    """oracle/tgn_oracle.py::tgn_memory_backward (float64 chain rule through GRUCell, the
    LastAggregator's pick and Time2Vec) against .grad of every parameter of the unmodified
    reference TGNMemory after (z * G).sum().backward() in the training loop of
    examples/linkproppred/tgn.py (tests/golden/make_golden_tgn.py::run_grad)."""
    from oracle.tgn_oracle import tgn_memory_backward
    z = np.load(path)
    p = _params(z)
    N, bs, rec = int(z['N']), int(z['bs']), int(z['record_from'])
    D, M, TD = z['x'].shape[1], p['memory_updater.weight_hh'].shape[1], p['time_enc.w.bias'].shape[0]
    mem = TGNMemoryOracle(N, D, M, TD, p)
    E, checked = len(z['src']), 0
    for b, lo in enumerate(range(0, E, bs)):
        hi = min(lo + bs, E)
        if b >= rec:
            n_id = z[f'b{b}_nid']
            zz, _ = mem.forward(n_id)
            assert np.abs(zz - z[f'b{b}_z']).max() <= 1e-5, b
            g = tgn_memory_backward(mem, n_id, z[f'b{b}_G'])
            names = [k.split('_g.', 1)[1] for k in z.files if k.startswith(f'b{b}_g.')]
            assert set(names) == set(g) and len(names) == 6
            for name in names:
                want = z[f'b{b}_g.{name}']
                assert g[name].shape == want.shape, name
                assert np.abs(g[name] - want).max() <= 2e-4 * max(1.0, np.abs(want).max()), (b, name)
            checked += 1
        mem.update_state(z['src'][lo:hi], z['dst'][lo:hi], z['t'][lo:hi], z['x'][lo:hi])
    assert checked >= 3


def test_graph_attention_embedding_backward_oracle_agrees_with_torch_autograd():
    """The embedding's convolution is third-party (parity UNPINNED, see above); its hand-derived
    backward is checked against autograd through the independent torch restatement, Time2Vec
    included (bias != 0 so that d bias is exercised)."""
    import math

    import torch
    from oracle.tgn_oracle import graph_attention_embedding, graph_attention_embedding_backward
    rng = np.random.default_rng(9)
    n, m, IN, HC, H, D, TD = 50, 500, 10, 12, 2, 4, 6
    p = {f'conv.{nm}.weight': rng.standard_normal((HC, IN)).astype(np.float32) * 0.4
         for nm in ('lin_query', 'lin_key', 'lin_value', 'lin_skip')}
    p.update({f'conv.{nm}.bias': rng.standard_normal(HC).astype(np.float32) * 0.1
              for nm in ('lin_query', 'lin_key', 'lin_value', 'lin_skip')})
    p['conv.lin_edge.weight'] = rng.standard_normal((HC, TD + D)).astype(np.float32) * 0.4
    p['time_enc.w.weight'] = (1 / 10 ** np.linspace(0, 3, TD)).astype(np.float32).reshape(TD, 1)
    p['time_enc.w.bias'] = rng.standard_normal(TD).astype(np.float32) * 0.3
    x = rng.standard_normal((n, IN)).astype(np.float32)
    ei = np.stack([rng.integers(0, n, m), rng.integers(0, n // 2, m)])
    ei[1, : m // 5] = 7  # a hub target
    lu, t = rng.integers(0, 50, n), rng.integers(0, 50, m)  # small deltas: float32 args stay exact
    msg = rng.standard_normal((m, D)).astype(np.float32)
    G = rng.standard_normal((n, HC)).astype(np.float32)

    tp = {k: torch.from_numpy(v).double().requires_grad_() for k, v in p.items()}
    xt = torch.from_numpy(x).double().requires_grad_()
    j, i = torch.from_numpy(ei[0]), torch.from_numpy(ei[1])
    rel = torch.from_numpy((lu[ei[0]] - t).astype(np.float64))
    enc = torch.cos(rel[:, None] * tp['time_enc.w.weight'].reshape(1, -1) + tp['time_enc.w.bias'])
    ea = torch.cat([enc, torch.from_numpy(msg).double()], 1)
    C = HC // H
    lin = lambda nm, v, bias=True: torch.nn.functional.linear(
        v, tp[f'conv.{nm}.weight'], tp[f'conv.{nm}.bias'] if bias else None)
    q, k, v = (lin(nm, xt).view(n, H, C) for nm in ('lin_query', 'lin_key', 'lin_value'))
    e = lin('lin_edge', ea, bias=False).view(-1, H, C)
    s = (q[i] * (k[j] + e)).sum(-1) / math.sqrt(C)
    mx = torch.full((n, H), float('-inf'), dtype=torch.float64).scatter_reduce(
        0, i[:, None].expand(-1, H), s.detach(), 'amax', include_self=True)
    ex = (s - mx[i]).exp()
    den = torch.zeros(n, H, dtype=torch.float64).index_add(0, i, ex) + 1e-16
    a = ex / den[i]
    out = torch.zeros(n, H, C, dtype=torch.float64).index_add(0, i, (v[j] + e) * a[:, :, None])
    out = out.view(n, HC) + lin('lin_skip', xt)
    fwd = graph_attention_embedding(p, H, x, lu, ei, t, msg)
    assert np.abs(out.detach().numpy() - fwd).max() <= 1e-5
    (out * torch.from_numpy(G).double()).sum().backward()

    g = graph_attention_embedding_backward(p, H, x, lu, ei, t, msg, G)
    assert set(g) == set(p) | {'x'}
    for name, want in [(k, tp[k].grad.numpy()) for k in p] + [('x', xt.grad.numpy())]:
        assert g[name].shape == want.shape, name
        # 1e-5: the oracle takes -sin of the float32-rounded Time2Vec argument (as the forward
        # rounds it), torch here of the float64 one
        assert np.abs(g[name] - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), name


# ---- TGN memory with the MeanAggregator (tgn.py:59-63) -------------------------------------------------
@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'tgnmean_*.npz'))),
                         ids=lambda p: os.path.basename(p)[8:-4])
def test_tgn_mean_aggregator_oracle_matches_reference(path):
    z = np.load(path)
    p = _params(z)
    N, bs, eval_from = int(z['N']), int(z['bs']), int(z['eval_from'])
    D, M, TD = z['x'].shape[1], p['memory_updater.weight_hh'].shape[1], p['time_enc.w.bias'].shape[0]
    assert int(z['mean']) == 1
    mem = TGNMemoryOracle(N, D, M, TD, p, aggregator='mean')
    E = len(z['src'])
    for b, lo in enumerate(range(0, E, bs)):
        hi = min(lo + bs, E)
        if b == eval_from:
            mem.train(False)
            assert np.abs(mem.memory - z['flush_memory']).max() <= 1e-5
            assert np.array_equal(mem.last_update, z['flush_last_update'])
        zz, lu = mem.forward(z[f'b{b}_nid'])
        assert np.abs(zz - z[f'b{b}_z']).max() <= 1e-5, b
        assert np.array_equal(lu, z[f'b{b}_lu']), b
        mem.update_state(z['src'][lo:hi], z['dst'][lo:hi], z['t'][lo:hi], z['x'][lo:hi])
    assert np.abs(mem.memory - z['final_memory']).max() <= 1e-5
    assert np.array_equal(mem.last_update, z['final_last_update'])


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'tgnmeangrad_*.npz'))),
                         ids=lambda p: os.path.basename(p)[12:-4])
def test_tgn_mean_aggregator_backward_oracle_matches_reference_autograd(path):
    from oracle.tgn_oracle import tgn_memory_backward
    z = np.load(path)
    p = _params(z)
    N, bs, rec = int(z['N']), int(z['bs']), int(z['record_from'])
    D, M, TD = z['x'].shape[1], p['memory_updater.weight_hh'].shape[1], p['time_enc.w.bias'].shape[0]
    mem = TGNMemoryOracle(N, D, M, TD, p, aggregator='mean')
    E, checked, multi = len(z['src']), 0, 0
    for b, lo in enumerate(range(0, E, bs)):
        hi = min(lo + bs, E)
        if b >= rec:
            n_id = z[f'b{b}_nid']
            zz, _ = mem.forward(n_id)
            assert np.abs(zz - z[f'b{b}_z']).max() <= 1e-5, b
            _, rows, _, weight = mem.aggregated_messages(n_id)
            multi += int((weight < 1).sum())
            g = tgn_memory_backward(mem, n_id, z[f'b{b}_G'])
            for name in [k.split('_g.', 1)[1] for k in z.files if k.startswith(f'b{b}_g.')]:
                want = z[f'b{b}_g.{name}']
                assert np.abs(g[name] - want).max() <= 2e-4 * max(1.0, np.abs(want).max()), (b, name)
            checked += 1
        mem.update_state(z['src'][lo:hi], z['dst'][lo:hi], z['t'][lo:hi], z['x'][lo:hi])
    assert checked >= 3 and multi > 0  # nodes with several messages were exercised


# ---- uniform sampler INCLUDING the reference's random.sample sub-sampling ------------------------------
@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'uniformrng_*.npz'))),
                         ids=lambda p: os.path.basename(p)[11:-4])
def test_uniform_oracle_reproduces_the_reference_random_sample_stream(path):
    """k below the degrees: get_nbrs keeps random.sample(candidates, k) per unique seed node
    (array_backend.py:147-153).  With random.seed fixed as the fixture generator fixed it, the oracle
    -- and the candidate-ordinal form the device path consumes -- reproduce every draw bit for bit."""
    import random

    from oracle.recency_oracle import (reference_rng_picks, uniform_candidates,
                                       uniform_sample_reference_rng)
    z = np.load(path)
    src, dst, t = z['src'], z['dst'], z['t']
    x = z['x'] if int(z['has_x']) else None
    bs, nn, directed = int(z['bs']), [int(v) for v in z['num_nbrs']], bool(int(z['directed']))
    subsampled = 0
    for ordinal_form in (False, True):
        random.seed(int(z['rng_seed']))
        for b, lo in enumerate(range(0, len(src), bs)):
            hi = min(lo + bs, len(src))
            e_hi = int(np.searchsorted(t, t[lo:hi].min() - 1, 'right'))
            seeds = np.concatenate([src[lo:hi], dst[lo:hi]])
            for h, k in enumerate(nn):
                if h:
                    seeds = nid.reshape(-1)
                if not ordinal_form:
                    nid, nt, nx = uniform_sample_reference_rng(src, dst, t, x, 0, e_hi, seeds, k, directed)
                else:  # what the device path does: counts -> host draws -> gather by ordinal
                    cand = uniform_candidates(src, dst, 0, e_hi, seeds, directed)
                    uniq = sorted(cand)
                    picks = reference_rng_picks([len(cand[v]) for v in uniq], k)
                    subsampled += sum(len(cand[v]) > k for v in uniq)
                    D = 0 if x is None else x.shape[1]
                    nid = np.full((len(seeds), k), -1, np.int32)
                    nt = np.zeros((len(seeds), k), np.int64)
                    nx = np.zeros((len(seeds), k, D), np.float32)
                    for i, v in enumerate(uniq):
                        rows = np.flatnonzero(seeds == v)
                        for j, p in enumerate(picks[i]):
                            if p >= 0:
                                e, nb = cand[v][p]
                                nid[rows, j], nt[rows, j] = nb, t[e]
                                if D:
                                    nx[rows, j] = x[e]
                assert np.array_equal(nid, z[f'b{b}_h{h}_nid']), (ordinal_form, b, h)
                assert np.array_equal(nt, z[f'b{b}_h{h}_nt'])
                assert np.array_equal(nx, z[f'b{b}_h{h}_nx'])
    assert subsampled > 20


def test_c_oracle_equals_the_unmodified_reference_on_the_config1_epoch():
    """BASELINE configs[0] at full size: the C ring oracle against what the unmodified reference
    put on every batch of the wiki-shaped epoch (tests/golden/make_golden_config1.py; seeds
    src + dst + the reference sampler's negatives, k=10, bs=200).  The stream lies outside the
    reference's int32 sort-key domain; on the batches the generator lists in `differs_from_ideal`
    (1 of 788) the expectation is the reference with its `.long()` fix."""
    import os

    from oracle.c_oracle import CRing, checksum_np
    from tests._golden import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, 'config1_wiki_epoch.npz'))
    E, N, D, bs, k = (int(z[n]) for n in ('E', 'N', 'D', 'bs', 'k'))
    rng = np.random.default_rng(int(z['seed']))
    src = rng.integers(0, 8227, E).astype(np.int32)
    dst = rng.integers(8227, N, E).astype(np.int32)
    t = np.sort(rng.integers(0, 2_678_374, E)).astype(np.int64)
    x = rng.standard_normal((E, D)).astype(np.float32)
    assert [checksum_np(v) for v in (src, dst, t, x)] == [int(v) for v in z['input_csum']]
    differs = {int(b): i for i, b in enumerate(z['differs_from_ideal'])}
    neg = z['neg']
    oracle = CRing(N, [k], D)
    for b, lo in enumerate(range(0, E, bs)):
        hi = min(lo + bs, E)
        seeds = np.concatenate([src[lo:hi], dst[lo:hi], neg[lo:hi]]).astype(np.int32)
        w = oracle.hook_call(seeds, np.concatenate([t[lo:hi]] * 3), src[lo:hi], dst[lo:hi],
                             t[lo:hi], x[lo:hi])[0]
        want = z['patched_csum'][differs[b]] if b in differs else z['csum'][b]
        assert [checksum_np(v) for v in w[2:]] == [int(v) for v in want], f'batch {b}'
        if f'b{b}_nid' in z.files and b not in differs:
            assert np.array_equal(w[2], z[f'b{b}_nid']) and np.array_equal(w[3], z[f'b{b}_nt'])


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'timeunit_*.npz'))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_ring_oracles_match_the_reference_under_time_window_batching(path):
    """Time-unit batches (tgm/data/loader.py:101-156) through the reference hook: uneven and empty
    windows.  The C and numpy ring oracles driven with the fixture's batch ranges."""
    from oracle.c_oracle import CRing
    from oracle.recency_oracle import RingOracle
    z = np.load(path)
    N, nn, directed = int(z['N']), [int(v) for v in z['num_nbrs']], bool(int(z['directed']))
    x = z['x'] if int(z['has_x']) else None
    D = 0 if x is None else x.shape[1]
    for oracle in (CRing(N, nn, D, directed), RingOracle(N, nn, D, directed)):
        for b in range(int(z['nb'])):
            lo, hi = int(z[f'b{b}_lo']), int(z[f'b{b}_hi'])
            s = np.concatenate([z['src'][lo:hi], z['dst'][lo:hi]]).astype(np.int32)
            q = np.concatenate([z['t'][lo:hi]] * 2)
            got = oracle.hook_call(s, q, z['src'][lo:hi], z['dst'][lo:hi], z['t'][lo:hi],
                                   None if x is None else x[lo:hi])
            for h in range(len(nn)):
                for u, name in zip(got[h][2:], ('nid', 'nt', 'nx')):
                    assert np.array_equal(u, z[f'b{b}_h{h}_{name}']), f'batch {b} hop {h} {name}'

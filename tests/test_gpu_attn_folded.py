"""The folded inference chain of TemporalAttention (csrc/attn_fold.cu) against the numpy oracle of
tgm/nn/modules/attention.py:58-128 and against the unfolded chain the backward pass keeps, over the
shapes its template instantiations split on; MergeLayer on both GEMM engines."""
import numpy as np
import pytest
import torch

from oracle import nn_oracle

pytestmark = pytest.mark.gpu

from tgm_b200 import _cabi  # noqa: E402
from tgm_b200.nn import TemporalAttention, Time2Vec  # noqa: E402
from tgm_b200.nn.attention import MergeLayer  # noqa: E402

DEV = torch.device('cuda', 0)
TOL = 1e-5


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _set(name, v):
    _cabi.check(_cabi.lib.tgm_set_option(name, v))


def _case(seed, S, k, nd, ed, td, H, pad_mode):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    att = TemporalAttention(H, nd, ed, td).to(DEV).eval()
    te = Time2Vec(td).to(DEV)
    with torch.no_grad():
        att.layer_norm.weight.uniform_(0.5, 1.5)
        att.layer_norm.bias.normal_()
        att.W_O.bias.normal_()
    node_x = rng.standard_normal((S, nd)).astype(np.float32)
    nbr_feat = rng.standard_normal((S, k, nd)).astype(np.float32)
    edge_feat = rng.standard_normal((S, k, ed)).astype(np.float32)
    seed_t = rng.integers(0, 2_678_373, S)
    nbr_t = np.sort(np.clip(seed_t[:, None] - rng.integers(1, 300_000, (S, k)), 0, None), 1)
    nbr_id = rng.integers(0, 9000, (S, k)).astype(np.int32)
    if pad_mode == 'ragged':  # left-padded slots, some seeds without any neighbour
        n_pad = rng.integers(0, k + 1, S)
        n_pad[: max(1, S // 10)] = k
        pad = np.arange(k)[None, :] < n_pad[:, None]
        nbr_id[pad], nbr_t[pad], edge_feat[pad] = -1, 0, 0.0
    p = {k_: v.detach().cpu().numpy() for k_, v in att.state_dict().items()}
    p.update({'time_encoder.' + k_: v.detach().cpu().numpy() for k_, v in te.state_dict().items()})
    want = nn_oracle.temporal_attention(
        p, '', H, node_x, nn_oracle._t2v(p, 'time_encoder.', np.zeros(S, np.int64)), edge_feat,
        nbr_feat, nn_oracle._t2v(p, 'time_encoder.', seed_t[:, None] - nbr_t), nbr_id != -1)
    return att, te, (node_x, nbr_feat, edge_feat, seed_t, nbr_t, nbr_id), want


SHAPES = [
    # S,   k,  nd,  ed,  td, H
    (700, 20, 1, 172, 100, 2),    # TGAT layer 1 on tgbl-wiki: qk built in the kernel
    (300, 20, 172, 172, 100, 2),  # TGAT layer 2: qk's x part from the product
    (257, 10, 3, 16, 100, 2),     # narrow edge features, odd out_dim (padding column)
    (129, 32, 4, 64, 32, 1),      # one head, k at the kernel's limit
    (65, 1, 40, 60, 7, 2),        # a single slot, short time encoding
    (64, 7, 100, 20, 128, 2),     # wide node, narrow edge, time_dim at the limit
    (33, 5, 5, 190, 65, 2),       # node_dim just past the in-kernel limit
]


@pytest.mark.parametrize('pad_mode', ['full', 'ragged'])
@pytest.mark.parametrize('shape', SHAPES, ids=lambda s: 'x'.join(map(str, s)))
def test_folded_chain_matches_the_oracle_and_the_unfolded_chain(shape, pad_mode):
    att, te, arrs, want = _case(3, *shape, pad_mode)
    args = [T(a) for a in arrs]
    outs = {}
    try:
        for folded, tc in ((1, 2), (1, 0), (0, 2)):
            _set(b'attn_folded', folded)
            _set(b'tc_linear', tc)
            with torch.no_grad():
                outs[(folded, tc)] = att.forward_fused(te, *args).cpu().numpy()
    finally:
        _set(b'attn_folded', 1)
        _set(b'tc_linear', 1)  # the default
    for key, got in outs.items():
        assert np.isfinite(got).all(), key
        assert np.abs(got - want).max() <= TOL, (key, float(np.abs(got - want).max()))


def test_folded_chain_reads_lazy_edge_rows_in_place():
    """edge features given as (table, row ids) -- the sampler's lazy form -- with -1 rows."""
    from tgm_b200.sampler import LazyEdgeRows
    S, k, nd, ed, td, H = 500, 20, 1, 172, 100, 2
    att, te, arrs, _ = _case(5, S, k, nd, ed, td, H, 'ragged')
    node_x, nbr_feat, _, seed_t, nbr_t, nbr_id = arrs
    rng = np.random.default_rng(9)
    table = rng.standard_normal((4000, ed)).astype(np.float32)
    rows = rng.integers(0, 4000, (S, k)).astype(np.int32)
    rows[nbr_id == -1] = -1
    dense = np.where((rows >= 0)[..., None], table[np.maximum(rows, 0)], 0.0).astype(np.float32)
    a = [T(node_x), T(nbr_feat)]
    b = [T(seed_t), T(nbr_t), T(nbr_id)]
    lazy = LazyEdgeRows(T(table), T(rows))
    with torch.no_grad():
        got_dense = att.forward_fused(te, *a, T(dense), *b)
        got_lazy = att.forward_fused(te, *a, lazy, *b)
    assert torch.equal(got_dense, got_lazy)


def test_folded_weights_follow_parameter_updates():
    att, te, arrs, _ = _case(7, 300, 20, 1, 172, 100, 2, 'full')
    args = [T(a) for a in arrs]
    with torch.no_grad():
        before = att.forward_fused(te, *args).clone()
        att.W_KV.weight.mul_(1.5)
        att.W_Q.weight.add_(0.01)
        te.w.bias.add_(0.1)
        after = att.forward_fused(te, *args)
    assert float((after - before).abs().max()) > 1e-3
    try:
        _set(b'attn_folded', 0)
        with torch.no_grad():
            unfolded = att.forward_fused(te, *args)
    finally:
        _set(b'attn_folded', 1)
    assert float((after - unfolded).abs().max()) <= TOL


@pytest.mark.parametrize('dims', [(102, 1, 172, 172, 12600), (272, 172, 172, 172, 600),
                                  (7, 3, 10, 5, 100), (102, 1, 172, 172, 1)],
                         ids=lambda d: 'x'.join(map(str, d)))
def test_merge_layer_on_both_gemm_engines(dims):
    in1, in2, hidden, out, S = dims
    torch.manual_seed(0)
    m = MergeLayer(in1, in2, hidden, out).to(DEV).eval()
    x1 = torch.randn(S, in1, device=DEV)
    x2 = torch.randn(S, in2, device=DEV)
    want = (torch.cat([x1, x2], 1).double() @ m.fc1.weight.double().T + m.fc1.bias.double()).relu()
    want = want @ m.fc2.weight.double().T + m.fc2.bias.double()
    try:
        for tc in (2, 0):
            _set(b'tc_linear', tc)
            with torch.no_grad():
                got = m(x1, x2)
            assert float((got.double() - want).abs().max()) <= TOL, tc
    finally:
        _set(b'tc_linear', 1)  # the default


@pytest.mark.parametrize('L,lazy', [(1, False), (2, False), (2, True), (3, False), (3, True)])
def test_tgat_runs_each_layer_once_over_all_hops(L, lazy):
    """TGAT.forward without gradients batches the hops of a layer into one attention + one merge
    call; with the folded chain switched off it walks the hops one by one (tgat.py:136-147).
    Same embeddings either way, dense or lazy edge features."""
    from tgm_b200.nn import TGAT
    from tgm_b200.sampler import LazyEdgeRows
    rng = np.random.default_rng(4)
    torch.manual_seed(4)
    N, nd, ed, td, emb, k, S0 = 500, 3, 16, 20, 24, 4, 30
    model = TGAT(nd, ed, td, emb, L, 2).to(DEV).eval()
    node_x = T(rng.standard_normal((N, nd)).astype(np.float32))
    table = T(rng.standard_normal((3000, ed)).astype(np.float32))
    seeds, times, nids, ex, nt = [], [], [], [], []
    cur = rng.integers(0, N, S0).astype(np.int32)
    cur_t = rng.integers(1000, 100_000, S0)
    for h in range(L):
        S = len(cur)
        nb = rng.integers(0, N, (S, k)).astype(np.int32)
        tt = np.sort(np.clip(cur_t[:, None] - rng.integers(1, 900, (S, k)), 0, None), 1)
        rows = rng.integers(0, 3000, (S, k)).astype(np.int32)
        pad = np.arange(k)[None, :] < rng.integers(0, k + 1, S)[:, None]
        nb[pad], tt[pad], rows[pad] = -1, 0, -1
        seeds.append(T(cur)), times.append(T(cur_t)), nids.append(T(nb)), nt.append(T(tt))
        lz = LazyEdgeRows(table, T(rows))
        ex.append(lz if lazy else lz.materialize().clone())
        cur, cur_t = nb.reshape(-1), np.repeat(cur_t, k)
    with torch.no_grad():
        batched = model(node_x, seeds, times, nids, ex, nt)
        try:
            _set(b'attn_folded', 0)
            per_hop = model(node_x, seeds, times, nids, ex, nt)
        finally:
            _set(b'attn_folded', 1)
    assert batched.shape == (S0, emb)
    assert float((batched - per_hop).abs().max()) <= TOL


def test_native_tgat_and_segment_calls_reject_bad_arguments():
    import ctypes
    from tgm_b200.nn import TGAT
    torch.manual_seed(0)
    model = TGAT(3, 16, 20, 24, 2, 2).to(DEV).eval()
    a = [m._handle(model.time_encoder, DEV) for m in model.attn]
    g = [m._handle(DEV) for m in model.merge_layers]
    h = ctypes.c_void_p()
    arr = lambda hs: (ctypes.c_void_p * len(hs))(*hs)  # noqa: E731
    lib = _cabi.lib
    assert lib.tgm_tgat_create(ctypes.byref(h), 5, arr(a), arr(g), 0) != 0       # too many layers
    assert lib.tgm_tgat_create(ctypes.byref(h), 2, arr(a), arr(g), -1) != 0      # no CPU form
    assert lib.tgm_tgat_create(ctypes.byref(h), 2, arr(a[::-1]), arr(g), 0) != 0  # dims do not chain
    assert lib.tgm_tgat_create(ctypes.byref(h), 2, arr(a), arr(g), 0) == 0 and h.value
    try:
        x = torch.zeros(10, 3, device=DEV)
        st = _cabi.current_stream(DEV)
        # NULL arrays; both / neither edge-feature form
        assert lib.tgm_tgat_forward(h, x.data_ptr(), 10, None, 4, None, None, None, None, None,
                                    None, 4, None, st) != 0
        assert 'tgm_tgat_forward' in _cabi.last_error()
        # segments that do not add up to S
        S, k = 6, 4
        nid = torch.zeros(S, k, dtype=torch.int32, device=DEV)
        t64 = torch.zeros(S, k, dtype=torch.int64, device=DEV)
        ef = torch.zeros(S, k, 16, device=DEV)
        out = torch.zeros(S, 24, device=DEV)
        nf = torch.zeros(S, k, 3, device=DEV)
        segs = (ctypes.c_void_p * 2)(ef.data_ptr(), ef.data_ptr())
        rows = (ctypes.c_int64 * 2)(2, 3)
        assert lib.tgm_attn_forward_segments(a[0], x.data_ptr(), nf.data_ptr(), segs, rows, 2,
                                             t64.data_ptr(), t64.data_ptr(), nid.data_ptr(), S, k,
                                             out.data_ptr(), st) != 0
        assert lib.tgm_attn_folded_covers(a[0], 33) == 0 and lib.tgm_attn_folded_covers(a[0], 32) == 1
    finally:
        lib.tgm_tgat_destroy(h)

"""Pin the CPU oracle: known answers of the reference's own unit tests + reference-generated
fixtures (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle.recency_oracle import PADDED_NODE_ID, RingOracle, masked_mean, stateless_sample
from tests._golden import Golden, assert_hop_equal, golden_files, golden_ids


@pytest.mark.parametrize('path', golden_files(), ids=golden_ids())
def test_ring_oracle_matches_reference_fixture(path):
    g = Golden(path)
    ring = RingOracle(g.N, g.num_nbrs, g.D, g.directed)
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

"""GPU parity of the hop-0 forms that never materialise the (S, k, D) feature block:
tgm_csr_sample_edges_ids / _mean and their host-buffer forms (include/tgm_b200.h), against the
reference-generated fixtures, the C oracle and the full-row path.  Also the store without a host
mirror of the timestamps (device bounds search, device order check) and the scan-built anchors."""
import numpy as np
import pytest
import torch

from oracle.c_oracle import CRing, masked_mean
from tests._golden import Golden, golden_files, golden_ids

pytestmark = pytest.mark.gpu

from tgm_b200 import RecencyCSR, _cabi  # noqa: E402
from tgm_b200.core.storage import DeviceCOOStorage, DGSliceTracker  # noqa: E402

DEV = 'cuda:0'


def _stream(seed, N, E, T, D, hot=0.0):
    rng = np.random.default_rng(seed)
    src, dst = rng.integers(0, N, E), rng.integers(0, N, E)
    if hot:
        src = np.where(rng.random(E) < hot, 1, src)
    t = np.sort(rng.integers(0, T, E))
    x = rng.standard_normal((E, D)).astype(np.float32) if D else None
    return src.astype(np.int32), dst.astype(np.int32), t.astype(np.int64), x


def _csr(src, dst, t, x, bs, directed=False, N=None):
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)  # noqa: E731
    store = DeviceCOOStorage.from_device_tensors(
        dev(src), dev(dst), dev(t), None if x is None else dev(x),
        int(max(src.max(), dst.max())) + 1 if N is None else N)
    return store, RecencyCSR(store, bs, directed=directed, colocate_x=True)


def _gather_rows(x, eid, D):
    """What a caller that owns edge_x does with the eid output."""
    out = np.zeros(eid.shape + (D,), np.float32)
    if D:
        m = eid >= 0
        out[m] = x[eid[m]]
    return out


def _ids_form_cases():
    """Every fixture whose hop-0 seeds are [src | dst] with B <= 32; directed adjacencies only in
    the search form (a dst endpoint owns no entry of its own there)."""
    cases, names = [], []
    for path, name in zip(golden_files(), golden_ids()):
        g = Golden(path)
        if g.neg is not None or max(g.num_nbrs) > 32:
            continue
        for search in (False, True):
            if g.directed and not search:
                continue
            cases.append((path, search))
            names.append(f"{name}-{'search' if search else 'anchors'}")
    return cases, names


@pytest.mark.parametrize('path,search', _ids_form_cases()[0], ids=_ids_form_cases()[1])
def test_ids_form_matches_reference_fixture(path, search):
    """(nid, t, eid) of every batch == the reference's nbr_nids / nbr_edge_time, and
    edge_x[eid] == its nbr_edge_x, for hop 0 of every fixture whose seeds are [src | dst]."""
    g = Golden(path)
    k, B = g.num_nbrs[0], max(g.num_nbrs)
    _, csr = _csr(g.src, g.dst, g.t, g.x, g.bs, g.directed, g.N)
    nid, nt, eid = (v.cpu().numpy() for v in csr.sample_edges_ids(0, g.E, k, B, search=search))
    rows = _gather_rows(g.x, eid, g.D)
    at = 0
    for b, lo, hi in g.batches():
        n = 2 * (hi - lo)
        want = g.expect(0, b, 0)
        assert np.array_equal(nid[at:at + n], want[2]), f'batch {b} nids'
        assert np.array_equal(nt[at:at + n], want[3]), f'batch {b} times'
        assert np.array_equal(rows[at:at + n], want[4]), f'batch {b} rows via eid'
        assert np.array_equal(eid[at:at + n] < 0, want[2] < 0)
        at += n


@pytest.mark.parametrize('cfg', [
    # N, E, T, D, bs, k, B, hot
    (2000, 60000, 900, 16, 200, 20, 20, 0.0),
    (300, 40000, 200, 4, 200, 5, 20, 0.05),      # k < B, > B pushes per node per batch
    (1000, 30000, 3000, 172, 200, 10, 10, 0.0),  # wiki-sized rows: 43 float4 columns > 32 lanes
    (500, 20011, 700, 8, 64, 32, 32, 0.2),       # ragged last batch, full-warp window
    (64, 3000, 100, 12, 333, 7, 7, 0.3),
], ids=['k20', 'k_lt_B_hot', 'D172', 'ragged_B32', 'D12'])
@pytest.mark.parametrize('search', [False, True], ids=['anchors', 'search'])
def test_fused_mean_is_bit_identical_to_sample_then_masked_mean(cfg, search):
    N, E, T, D, bs, k, B, hot = cfg
    src, dst, t, x = _stream(11, N, E, T, D, hot)
    _, csr = _csr(src, dst, t, x, bs, N=N)
    nid, nt, nx = csr.sample_edges(0, E, k, B)
    want = masked_mean(nx.cpu().numpy(), nid.cpu().numpy())  # C oracle of graphmixer.py:131-135
    got = csr.sample_edges_mean(0, E, k, B, search=search)
    assert torch.equal(got[0], nid) and torch.equal(got[1], nt)
    assert np.array_equal(got[2].cpu().numpy(), want)
    only = csr.sample_edges_mean(0, E, k, B, search=search, with_ids=False)
    assert only[0] is None and np.array_equal(only[2].cpu().numpy(), want)
    # a window in the middle of the stream, on a batch boundary
    lo, hi = 3 * bs, min(E, 41 * bs)
    sub = csr.sample_edges_mean(lo, hi, k, B, search=search)
    a, b = 2 * lo, 2 * hi
    assert torch.equal(sub[0], nid[a:b]) and np.array_equal(sub[2].cpu().numpy(), want[a:b])


def test_ids_and_mean_against_the_c_oracle_batch_by_batch():
    """Independent of the full-row kernel: the ring oracle driven batch by batch."""
    N, E, T, D, bs, k = 700, 24000, 500, 8, 150, 9
    src, dst, t, x = _stream(5, N, E, T, D, 0.1)
    _, csr = _csr(src, dst, t, x, bs, N=N)
    nid, nt, eid = (v.cpu().numpy() for v in csr.sample_edges_ids(0, E, k, k, search=True))
    mean = csr.sample_edges_mean(0, E, k, k, with_ids=False)[2].cpu().numpy()
    rows = _gather_rows(x, eid, D)
    oracle = CRing(N, [k], D)
    for lo in range(0, E, bs):
        hi = min(lo + bs, E)
        s = np.concatenate([src[lo:hi], dst[lo:hi]])
        q = np.concatenate([t[lo:hi], t[lo:hi]])
        w = oracle.hook_call(s, q, src[lo:hi], dst[lo:hi], t[lo:hi], x[lo:hi])[0]
        a, b = 2 * lo, 2 * hi
        assert np.array_equal(nid[a:b], w[2]) and np.array_equal(nt[a:b], w[3])
        assert np.array_equal(rows[a:b], w[4])
        assert np.array_equal(mean[a:b], masked_mean(w[4], w[2]))


def test_host_forms_move_only_what_the_host_lacks():
    N, E, D, bs, k = 4000, 40_000, 8, 100, 12
    src, dst, t, x = _stream(3, N, E, 1500, D)
    _, csr = _csr(src, dst, t, x, bs, N=N)
    lo, hi = 20_000, 26_000
    n = 2 * (hi - lo)
    nid, nt, nx = (v.cpu() for v in csr.sample_edges(lo, hi, k, k))
    want_mean = torch.from_numpy(masked_mean(nx.numpy(), nid.numpy()))
    host_in = tuple(torch.from_numpy(np.ascontiguousarray(a[lo:hi])).pin_memory()
                    for a in (src, dst, t))
    for slab in (host_in, (None, None, None)):
        out = (torch.full((n, k), -7, dtype=torch.int32).pin_memory(),
               torch.full((n, k), -7, dtype=torch.int64).pin_memory(),
               torch.full((n, k), -7, dtype=torch.int32).pin_memory())
        csr.sample_edges_host_ids(lo, hi, k, k, slab, out)
        torch.cuda.synchronize()
        assert torch.equal(out[0], nid) and torch.equal(out[1], nt)
        assert np.array_equal(_gather_rows(x, out[2].numpy(), D), nx.numpy())
        mo = (torch.empty((n, k), dtype=torch.int32).pin_memory(),
              torch.empty((n, k), dtype=torch.int64).pin_memory(), torch.empty((n, D)).pin_memory())
        csr.sample_edges_host_mean(lo, hi, k, k, slab, mo)
        torch.cuda.synchronize()
        assert torch.equal(mo[0], nid) and torch.equal(mo[1], nt) and torch.equal(mo[2], want_mean)
        m2 = (None, None, torch.empty((n, D)).pin_memory())
        csr.sample_edges_host_mean(lo, hi, k, k, slab, m2, slot=1)
        torch.cuda.synchronize()
        assert torch.equal(m2[2], want_mean)
    # the search form really consumes the uploaded slab: upload different seeds, get their answers
    perm = torch.from_numpy(np.ascontiguousarray(src[lo:hi][::-1])).pin_memory()
    out = tuple(torch.empty((n, k), dtype=d).pin_memory()
                for d in (torch.int32, torch.int64, torch.int32))
    csr.sample_edges_host_ids(lo, hi, k, k, (perm, host_in[1], host_in[2]), out)
    torch.cuda.synchronize()
    seeds = torch.from_numpy(np.concatenate(
        [np.concatenate([perm.numpy()[a - lo:a - lo + bs], dst[a:a + bs]])
         for a in range(lo, hi, bs)])).to(DEV)
    tq = torch.from_numpy(np.concatenate(
        [np.concatenate([t[a:a + bs], t[a:a + bs]]) for a in range(lo, hi, bs)])).to(DEV)
    cut = torch.arange(lo, hi, bs, device=DEV, dtype=torch.int64)
    ref = csr.sample(seeds, tq, cut, k, k, cut_group=2 * bs)
    assert torch.equal(out[0].to(DEV), ref[0]) and torch.equal(out[1].to(DEV), ref[1])
    csr.sample_edges_host_ids(lo, hi, k, k, host_in, out)  # restore the slab
    torch.cuda.synchronize()
    with pytest.raises(_cabi.TGMNativeError, match='bad slot'):
        csr.sample_edges_host_ids(lo, hi, k, k, host_in, out, slot=99)


def test_argument_errors_of_the_fused_forms():
    src, dst, t, x = _stream(1, 50, 2000, 100, 6)  # D % 4 != 0
    _, csr = _csr(src, dst, t, x, 50, N=50)
    with pytest.raises(_cabi.TGMNativeError, match='D % 4 == 0'):
        csr.sample_edges_mean(0, 2000, 5, 5)
    with pytest.raises(_cabi.TGMNativeError, match='batch boundary'):
        csr.sample_edges_ids(7, 2000, 5, 5)
    with pytest.raises(_cabi.TGMNativeError, match=r'B must be in \[1, 32\]'):
        csr.sample_edges_ids(0, 2000, 5, 40)
    nid, nt, eid = csr.sample_edges_ids(0, 0, 5, 5)  # empty window: nothing launched
    assert nid.shape == (0, 5)


# ---- store without a host mirror ---------------------------------------------------------------
def test_adopted_device_stream_keeps_no_host_mirror_and_searches_bounds_on_device():
    src, dst, t, x = _stream(2, 100, 5000, 300, 0)
    dev = lambda a: torch.from_numpy(a).to(DEV)  # noqa: E731
    store = DeviceCOOStorage.from_device_tensors(dev(src), dev(dst), dev(t), None, 100)
    assert store._time_np_cache is None
    for lo_t, hi_t in [(None, None), (0, 0), (17, 17), (17, 250), (None, 123), (299, None),
                       (300, 400), (-5, 2)]:
        for idx in [(None, None), (100, 4000), (4000, 100)]:
            s = DGSliceTracker(lo_t, hi_t, *idx)
            lb = 0 if lo_t is None else int(np.searchsorted(t, lo_t, 'left'))
            ub = len(t) if hi_t is None else int(np.searchsorted(t, hi_t, 'right'))
            cl, ch = idx[0] or 0, idx[1] or len(t)
            want = (max(cl, min(ch, lb)), max(cl, min(ch, ub)))
            assert store._event_bounds(s) == want, (lo_t, hi_t, idx)
    assert store._time_np_cache is None  # bounds never pulled the timestamps to the host
    assert store.get_start_time(DGSliceTracker()) == int(t[0])
    assert store.get_end_time(DGSliceTracker(end_idx=1234)) == int(t[1233])
    assert store.get_num_timestamps(DGSliceTracker()) == len(np.unique(t))  # this one reads back
    bad = t.copy()
    bad[2500] = bad[-1] + 5  # one spike in the middle of the stream
    with pytest.raises(_cabi.TGMNativeError, match='non-decreasing'):
        DeviceCOOStorage.from_device_tensors(dev(src), dev(dst), dev(bad), None, 100)


@pytest.mark.parametrize('bs', [1, 7, 200])
def test_scan_built_anchors_equal_a_search(bs):
    """The anchor of (edge, endpoint) = first entry of the endpoint that belongs to the edge's own
    batch or later; the build derives it from run heads + a max-scan, the general kernel from a
    binary search per seed: both must sample the same windows."""
    N, E = 150, 20000
    src, dst, t, x = _stream(9, N, E, 400, 4, 0.2)
    _, csr = _csr(src, dst, t, x, bs, N=N)
    a = csr.sample_edges(0, E, 6, 11)
    seeds = torch.from_numpy(np.concatenate(
        [np.concatenate([src[i:i + bs], dst[i:i + bs]]) for i in range(0, E, bs)])).to(DEV)
    tq = torch.from_numpy(np.concatenate(
        [np.concatenate([t[i:i + bs], t[i:i + bs]]) for i in range(0, E, bs)])).to(DEV)
    starts = torch.arange(0, E, bs, device=DEV, dtype=torch.int64)
    counts = torch.full((len(starts),), 2 * bs, device=DEV, dtype=torch.int64)
    if E % bs:
        counts[-1] = 2 * (E % bs)
    cut = torch.repeat_interleave(starts, counts)
    b = csr.sample(seeds, tq, cut, 6, 11)
    for u, v in zip(a, b):
        assert torch.equal(u, v)

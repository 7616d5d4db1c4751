"""The `DGStorageBase` contract (tgm/core/_storage/base.py:20-118) getter by getter against what the
reference's own backend returned on the same inputs (tests/golden/make_golden_storage.py ran
`DGStorageArrayBackend` unmodified): 4 graphs x 12 slices.  The metadata getters are checked on any
host; the getters that serve edge data need the B200 (`-m gpu`) -- the reference's
test_storage_impl.py cannot run there (no reference tree on the GPU box), so its subject matter is
pinned this way."""
import json
import os

import numpy as np
import pytest
import torch

from tests._golden import GOLDEN_DIR
from tgm_b200 import DGData
from tgm_b200.core.storage import DeviceCOOStorage, DGSliceTracker

Z = np.load(os.path.join(GOLDEN_DIR, 'storage_contract.npz'))
CASES = sorted({k.split('/')[0] for k in Z.files})
NSLICES = 12


def _js(key):
    return json.loads(bytes(Z[key]).decode())


def _store(case, device):
    import warnings
    kw = {k.split('/in_')[1]: torch.from_numpy(Z[k]) for k in Z.files if k.startswith(case + '/in_')}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        data = DGData.from_raw(time_delta='s', **kw)
    return DeviceCOOStorage(data, device=device)


@pytest.mark.parametrize('case', CASES)
def test_metadata_getters_match_the_reference_backend(case):
    st = _store(case, None)  # metadata-only store: no device needed
    dims = _js(case + '/dims')
    assert st.get_static_node_x_dim() == dims['static_node_x_dim']
    assert st.get_node_x_dim() == dims['node_x_dim'] and st.get_node_y_dim() == dims['node_y_dim']
    assert st.get_edge_x_dim() == dims['edge_x_dim']
    assert (st.get_static_node_x() is not None) == dims['has_static']
    assert (st.get_node_type() is not None) == dims['has_node_type']
    for si in range(NSLICES):
        pre = f'{case}/s{si}/'
        m = _js(pre + 'meta')
        s = DGSliceTracker(*m['slice'])
        assert st.get_start_time(s) == m['start_time'], (si, 'start_time')
        assert st.get_end_time(s) == m['end_time'], (si, 'end_time')
        assert st.get_num_events(s) == m['num_events'], (si, 'num_events')
        assert st.get_num_timestamps(s) == m['num_timestamps'], (si, 'num_timestamps')
        assert sorted(st.get_nodes(s)) == m['nodes'], (si, 'nodes')
        for tag, (ids, tt) in (('node_events', st.get_node_events(s)), ('node_labels', st.get_node_labels(s))):
            want = Z[pre + tag]
            assert np.array_equal(ids.numpy().astype(np.int64), want[0]), (si, tag)
            assert np.array_equal(tt.numpy().astype(np.int64), want[1]), (si, tag)
        for tag, sp in (('node_x', st.get_node_x(s)), ('node_y', st.get_node_y(s))):
            assert (sp is None) == m[tag + '_none'], (si, tag)
            if sp is not None:
                sp = sp.coalesce()
                assert np.array_equal(sp.indices().numpy(), Z[pre + tag + '_idx']), (si, tag)
                assert np.array_equal(sp.values().numpy(), Z[pre + tag + '_val']), (si, tag)
                assert list(sp.shape) == Z[pre + tag + '_shape'].tolist(), (si, tag)


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_edge_getters_match_the_reference_backend(case):
    st = _store(case, 'cuda:0')
    for si in range(NSLICES):
        pre = f'{case}/s{si}/'
        m = _js(pre + 'meta')
        s = DGSliceTracker(*m['slice'])
        src, dst, t = st.get_edges(s)
        assert src.dtype == torch.int32 and dst.dtype == torch.int32 and t.dtype == torch.int64
        assert src.is_cuda and src.is_contiguous()
        want = Z[pre + 'edges']
        for got, w in zip((src, dst, t), want):
            assert np.array_equal(got.cpu().numpy().astype(np.int64), w), (si, 'edges')
        ex = st.get_edge_x(s)
        assert (ex is None) == m['edge_x_none'], (si, 'edge_x')
        if ex is not None:
            assert np.array_equal(ex.cpu().numpy(), Z[pre + 'edge_x'])
        et = st.get_edge_type(s)
        assert (et is None) == m['edge_type_none'], (si, 'edge_type')
        if et is not None:
            assert np.array_equal(et.cpu().numpy(), Z[pre + 'edge_type'])
        if pre + 'nbr_seeds' in Z.files:  # get_nbrs with k >= every degree: the reference's rows
            seeds = torch.from_numpy(Z[pre + 'nbr_seeds'])
            for directed, tag in ((False, 'nbrs'), (True, 'nbrs_dir')):
                nid, nt, nx = st.get_nbrs(seeds, 64, s, directed)
                assert np.array_equal(nid.cpu().numpy(), Z[pre + tag + '_nid']), (si, tag)
                assert np.array_equal(nt.cpu().numpy(), Z[pre + tag + '_t']), (si, tag)
                assert np.array_equal(nx.cpu().numpy().reshape(Z[pre + tag + '_x'].shape),
                                      Z[pre + tag + '_x']), (si, tag)

"""GPU parity of the uniform sampler's reference-exact mode (NeighborSamplerHook(reference_rng=True):
tgm_csr_candidate_counts + the reference's own random.sample calls on the host +
tgm_csr_gather_picks): with `random.seed` fixed as the fixture generator fixed it, every output of
every hop of every batch equals the unmodified reference's, sub-sampled rows included.

"""
import glob
import os
import random

import numpy as np
import pytest
import torch

from tests._golden import GOLDEN_DIR

pytestmark = pytest.mark.gpu

from tgm_b200 import (DGData, DGDataLoader, DGraph, HookManager,  # noqa: E402
                      NeighborSamplerHook)

DEV = 'cuda:0'


def _graph(src, dst, t, x):
    ei = torch.from_numpy(np.stack([src, dst], 1).astype(np.int32))
    return DGraph(DGData.from_raw(torch.from_numpy(np.asarray(t, np.int64)), ei,
                                  None if x is None else torch.from_numpy(x)), device=DEV)


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(GOLDEN_DIR, 'uniform*.npz'))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_reference_rng_mode_matches_reference_fixture(path):
    """uniformrng_*: k below the degrees (random.sample on most queries); uniform_*: k above every
    degree (the exact mode must agree with the reference there as well)."""
    z = np.load(path)
    x = z['x'] if int(z['has_x']) else None
    dg = _graph(z['src'], z['dst'], z['t'], x)
    nn = [int(v) for v in z['num_nbrs']]
    hm = HookManager(keys=['g'])
    hm.register('g', NeighborSamplerHook(num_nbrs=nn, seed_nodes_keys=['edge_src', 'edge_dst'],
                                         seed_times_keys=['edge_time', 'edge_time'],
                                         directed=bool(int(z['directed'])), reference_rng=True))
    random.seed(int(z['rng_seed']))
    with hm.activate('g'):
        nb = 0
        for b, batch in enumerate(DGDataLoader(dg, batch_size=int(z['bs']), hook_manager=hm)):
            nb += 1
            for h in range(len(nn)):
                for name, got in (('seed', batch.seed_nids[h]), ('tq', batch.seed_times[h]),
                                  ('nid', batch.nbr_nids[h]), ('nt', batch.nbr_edge_time[h]),
                                  ('nx', batch.nbr_edge_x[h])):
                    want = z[f'b{b}_h{h}_{name}']
                    g = got.cpu().numpy()
                    assert g.dtype == want.dtype and g.shape == want.shape, (b, h, name)
                    assert np.array_equal(g, want), (b, h, name)
    assert nb == -(-len(z['src']) // int(z['bs']))

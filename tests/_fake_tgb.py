"""A stand-in for py-tgb's NegativeEdgeSampler (absent offline): a deterministic candidate list per
positive edge.  Used by tests/golden/make_golden_tgbneg.py (installed as the `tgb` package the
UNMODIFIED reference hook imports) and, injected as `neg_sampler=`, by the tests of our hook."""
import numpy as np


class FakeNegativeEdgeSampler:
    def __init__(self, dataset_name=None, num_nodes=50, **_):
        self.num_nodes = num_nodes
        self.loaded = None

    def load_eval_set(self, fname, split_mode):
        self.loaded = (fname, split_mode)

    def query_batch(self, src, dst, t, *edge_type, split_mode='val'):
        if split_mode not in ('val', 'test'):
            raise ValueError(split_mode)
        src, dst, t = (np.asarray(a.cpu() if hasattr(a, 'cpu') else a).astype(np.int64)
                       for a in (src, dst, t))
        et = np.asarray(edge_type[0].cpu()).astype(np.int64) if edge_type else np.zeros_like(src)
        out = []
        for s, d, tt, e in zip(src, dst, t, et):
            n = 1 + int((s + 2 * d + tt + e) % 5)
            out.append([int((7 * s + 3 * d + tt + 11 * j + 5 * e) % self.num_nodes)
                        for j in range(n)])
        return out

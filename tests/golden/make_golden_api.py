"""Snapshot of the reference's public signatures on the hot path (constructor / method parameter
names, order, kinds and defaults), taken from the UNMODIFIED reference:
    python tests/golden/make_golden_api.py   ->   tests/golden/api_signatures.json
tests/test_host_logic.py checks that the drop-in accepts every one of them (extra optional
parameters may follow)."""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shim import import_reference  # noqa: E402

import_reference()

# dotted path below the package root, identical in tgm and tgm_b200
NAMES = [
    'DGraph', 'DGraph.slice_time', 'DGraph.slice_events', 'DGraph.materialize', 'DGraph.to',
    'data.DGData.from_raw', 'data.DGDataLoader',
    'hooks.HookManager', 'hooks.HookManager.register', 'hooks.HookManager.register_shared',
    'hooks.HookManager.activate', 'hooks.HookManager.execute_active_hooks',
    'hooks.HookManager.reset_state', 'hooks.HookManager.validate_requirement',
    'hooks.RecencyNeighborHook', 'hooks.NeighborSamplerHook', 'hooks.DeduplicationHook',
    'hooks.RandomNegativeEdgeSamplerHook', 'hooks.StatelessHook', 'hooks.StatefulHook',
    'hooks.SeedableHook', 'hooks.BaseDGHook.add_batch_attribute',
    'nn.TGAT', 'nn.TGAT.forward', 'nn.DyGFormer', 'nn.DyGFormer.forward', 'nn.Time2Vec',
    'nn.TemporalAttention', 'nn.TemporalAttention.forward',
    'nn.encoder.tgn.TGNMemory', 'nn.encoder.tgn.TGNMemory.forward',
    'nn.encoder.tgn.TGNMemory.update_state', 'nn.encoder.tgn.TGNMemory.reset_state',
    'nn.encoder.tgn.TGNMemory.detach', 'nn.encoder.tgn.IdentityMessage',
    'nn.encoder.tgn.GraphAttentionEmbedding', 'nn.encoder.tgn.GraphAttentionEmbedding.forward',
]


from _api_sig import describe, resolve  # noqa: E402


def main():
    import tgm
    snap = {n: describe(resolve(tgm, n)) for n in NAMES}
    with open(os.path.join(HERE, 'api_signatures.json'), 'w') as f:
        json.dump(snap, f, indent=1, sort_keys=True)
    print(len(snap), 'signatures')


if __name__ == '__main__':
    main()

"""Aggregation modules of the hot path (tgm/nn of the reference, the parts SURVEY.md section 8
puts on the path): `torch.nn.Module`s with the reference's parameter names and shapes, so
`state_dict`s interchange, executing on the CUDA library (include/tgm_b200.h); differentiable with
dropout 0 (`tgm_attn_backward`, `tgm_dyg_backward`, `tgm_tgn_backward`, `tgm_gae_backward`)."""
from .attention import MergeLayer, TemporalAttention, Time2Vec, masked_mean
from .dygformer import DyGFormer
from .tgat import TGAT
from .tgn import (GraphAttentionEmbedding, IdentityMessage, LastAggregator, MeanAggregator,
                  TGNMemory, TransformerConv)

__all__ = ['TemporalAttention', 'Time2Vec', 'MergeLayer', 'TGAT', 'DyGFormer', 'masked_mean',
           'TGNMemory', 'IdentityMessage', 'LastAggregator', 'MeanAggregator', 'GraphAttentionEmbedding',
           'TransformerConv']


def _reference_import_paths() -> None:
    """`tgm.nn.encoder.{tgat,dygformer,tgn}` and `tgm.nn.modules.{attention,time_encoding}` (the
    reference's module layout) resolve to the modules of this package: aliases in sys.modules, not
    files of re-exports."""
    import sys
    import types
    from . import attention, dygformer, tgat, tgn
    base = __name__
    enc = types.ModuleType(f'{base}.encoder')
    enc.__path__ = []  # a package: `import tgm_b200.nn.encoder.tgat` finds the entry below
    enc.__all__ = ['DyGFormer', 'TGAT', 'GraphAttentionEmbedding', 'IdentityMessage',
                   'LastAggregator', 'MeanAggregator', 'TGNMemory']
    for name in enc.__all__:
        setattr(enc, name, globals()[name])
    mod = types.ModuleType(f'{base}.modules')
    mod.__path__ = []
    mod.__all__ = ['TemporalAttention', 'Time2Vec']
    mod.TemporalAttention, mod.Time2Vec = TemporalAttention, Time2Vec
    for pkg, subs in ((enc, {'tgat': tgat, 'dygformer': dygformer, 'tgn': tgn}),
                      (mod, {'attention': attention, 'time_encoding': attention})):
        sys.modules[pkg.__name__] = pkg
        for sub, target in subs.items():
            sys.modules[f'{pkg.__name__}.{sub}'] = target
            setattr(pkg, sub, target)
    globals()['encoder'], globals()['modules'] = enc, mod


_reference_import_paths()

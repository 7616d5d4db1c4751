"""DGBatch: what the loader yields and hooks decorate.  Field names, order and defaults follow
tgm/core/batch.py:32-46; hooks attach further attributes with setattr."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

from torch import Tensor


@dataclass
class DGBatch:
    edge_src: Tensor   # int32 (E_b,)
    edge_dst: Tensor   # int32 (E_b,)
    edge_time: Tensor  # int64 (E_b,)
    edge_x: Optional[Tensor] = None     # float32 (E_b, D)
    edge_type: Optional[Tensor] = None

    node_x_time: Optional[Tensor] = None
    node_x_nids: Optional[Tensor] = None
    node_x: Optional[Tensor] = None

    node_y_time: Optional[Tensor] = None
    node_y_nids: Optional[Tensor] = None
    node_y: Optional[Tensor] = None

    def __str__(self) -> str:
        def show(v) -> str:
            if isinstance(v, Tensor):
                return str(list(v.shape))
            if isinstance(v, (list, tuple)):
                return f'{type(v).__name__}({"|".join(sorted({show(u) for u in v}))} x{len(v)})'
            return type(v).__name__
        return 'DGBatch(' + ', '.join(f'{k} = {show(v)}' for k, v in vars(self).items()
                                       if not k.startswith('_')) + ')'

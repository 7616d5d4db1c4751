"""`tgm.nn.modules.attention` import path."""
from tgm_b200.nn.attention import TemporalAttention  # noqa: F401

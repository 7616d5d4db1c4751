"""`tgm.nn.modules.time_encoding` import path."""
from tgm_b200.nn.attention import Time2Vec  # noqa: F401

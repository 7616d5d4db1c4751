"""Constants shared with the reference API (tgm/constants.py)."""
from typing import Final

PADDED_NODE_ID: Final[int] = -1  # sentinel id of padded neighbour slots (tgm/constants.py:3)
PADDED_TIME: Final[int] = 0      # timestamp written to padded slots (recency.py:318)

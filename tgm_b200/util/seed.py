"""`tgm.util.seed.seed_everything` (tgm/util/seed.py:11-25): one seed for Python, numpy and torch
(CPU and every CUDA device) -- the examples call it first thing."""
import random

import numpy as np
import torch


def seed_everything(seed: int) -> None:
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)

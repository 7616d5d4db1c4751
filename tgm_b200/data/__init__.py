from .dg_data import DGData
from .loader import DGDataLoader

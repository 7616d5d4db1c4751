"""Small utilities with the reference's import paths (tgm/util)."""
from .seed import seed_everything

__all__ = ['seed_everything']

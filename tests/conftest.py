import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'slow: larger CPU-side case')


def pytest_addoption(parser):
    parser.addoption('--default-window-batches', type=int, default=0,
                     help='experiment: run the suite with RecencyNeighborHook(window_batches=N) as '
                          'the default (DESIGN.md section 8, item 4), to see whether the windowed '
                          'pre-sampling mode could become the default for device stores')


@pytest.fixture(autouse=True, scope='session')
def _default_window_batches(request):
    n = request.config.getoption('--default-window-batches', default=0)
    if not n:
        yield
        return
    from tgm_b200.hooks.recency import RecencyNeighborHook
    init = RecencyNeighborHook.__init__

    def patched(self, *args, **kwargs):
        if len(args) < 7:  # window_batches is the 7th positional parameter
            kwargs.setdefault('window_batches', n)
        init(self, *args, **kwargs)

    RecencyNeighborHook.__init__ = patched
    yield
    RecencyNeighborHook.__init__ = init


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)

"""pytest plugin (test infrastructure): imports the UNMODIFIED reference `tgm` package (with the
torch_geometric stub of tests/golden/_ref_shim.py) and registers this repo's storage backend in the
reference's own registry BEFORE the reference's test modules are collected, so that
`@pytest.fixture(params=DGStorageBackends.values())` (test/unit/test_core/test_storage_impl.py:23-25)
parametrises the reference's storage contract suite over it.
Loaded with `-p tests._reference_backend_plugin` (tests/test_reference_suite_differential.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests.golden._ref_shim import import_reference  # noqa: E402

import_reference()

from tgm_b200 import reference_plugin  # noqa: E402

# registered next to the array backend, not as the default: the suite asks for each by fixture
reference_plugin.install(make_default=False)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dbs():
    import json

    import numpy as np

    z = np.load(os.path.join(GOLDEN, "reference_dbs.npz"))
    alias = json.loads(str(z["alias_json"]))

    class G:
        headers = json.loads(str(z["headers_json"]))

        def __getitem__(self, k):
            return z[alias.get(k, k)]

        def __contains__(self, k):
            return alias.get(k, k) in z.files

    return G()


@pytest.fixture(scope="session")
def golden_static():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "reference_static_methods.npz"))

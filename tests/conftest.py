import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run by the driver with -m gpu)')
    config.addinivalue_line('markers', 'reference: needs /root/reference (build container only)')


@pytest.fixture(scope='session')
def golden():
    import numpy as np
    out = {}
    for name in ('layers', 'embeddings', 'models', 'layers2', 'models2'):
        with np.load(os.path.join(ROOT, 'tests', 'golden', f'{name}.npz')) as z:
            out.update({k: z[k] for k in z.files})
    return out

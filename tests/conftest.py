import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')

# the library caches its experiment switches at first use; tests flip them between calls (monkeypatch.setenv)
os.environ.setdefault('CNF_LIVE_ENV', '1')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def golden_flow_names():
    return sorted(f[len('flow_'):-len('.npz')] for f in os.listdir(GOLDEN) if f.startswith('flow_'))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def oracle_params_from_golden(g, dtype=np.float64):
    """Rebuild the oracle's parameter container from a golden flat vector."""
    import flow_oracle as orc
    K, L = int(g['K']), int(g['L'])
    hidden = [int(h) for h in g['hidden']]
    like = orc.init_params(K, L, hidden, bool(g['scale']), bool(g['shift']))
    for l, lay in enumerate(like):
        lay['perm'] = g['perms'][l].astype(np.int64) if int(g['random_flip']) else None
    return orc.unflatten(g['flat'].astype(dtype), like)


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')

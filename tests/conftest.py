import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


def pytest_sessionstart(session):
    """A fresh checkout has no libpixtrack_b200.so (built artefacts are git-ignored): build it once (nvcc cross-compiles
    without a GPU; the digest check makes this a no-op when the library is current)."""
    try:
        from pixtrack_b200 import build as b
        b.build()
    except Exception as e:  # noqa: BLE001 - the tests that need the library then fail with the real reason
        print(f'[conftest] library build failed: {e}', file=sys.stderr)


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)

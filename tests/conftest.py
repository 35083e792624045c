import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore", message="Covariance of the parameters could not be estimated")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return load


def clip_from_fixture(fix):
    """Regenerate the fixture's synthetic clip (bit-identical by construction)."""
    from respmon_b200 import synth
    W, H, T, seed = (int(v) for v in fix["spec"][:4])
    spec = synth.clip_spec(seed, W, H, T)
    assert [spec.x0, spec.y0, spec.w0, spec.h0] == [int(v) for v in fix["spec"][4:8]]
    return spec, synth.make_clip(spec)

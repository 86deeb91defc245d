import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def clip58():
    import numpy as np
    d = np.load(os.path.join(GOLDEN, "clip58.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def camera():
    from globalegomocap_b200 import synthetic as syn
    return syn.load_camera()


@pytest.fixture(scope="session")
def vae_weights(clip58):
    """The two random-init checkpoints the goldens were generated with
    (tests/golden/make_golden.py Harness: seeds 11/12, perturbed BN, mean-pose bias)."""
    from globalegomocap_b200 import synthetic as syn
    bias = syn.mean_pose_bias(clip58)
    return (syn.make_vae_state_dict(11, perturb_bn=True, pose_bias=bias),
            syn.make_vae_state_dict(12, perturb_bn=True, pose_bias=bias))


@pytest.fixture(scope="session")
def vae_weights_g2(clip58):
    """Checkpoints of tests/golden/traces_g2.npz (same seeds, gain=2)."""
    from globalegomocap_b200 import synthetic as syn
    bias = syn.mean_pose_bias(clip58)
    return (syn.make_vae_state_dict(11, perturb_bn=True, pose_bias=bias, gain=2.0),
            syn.make_vae_state_dict(12, perturb_bn=True, pose_bias=bias, gain=2.0))

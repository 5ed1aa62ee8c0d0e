import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "smart-vocoder_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def base_cfg():
    import json
    return json.load(open(os.path.join(ROOT, "configs", "iitp_base.json")))


@pytest.fixture(scope="session")
def base_dims(base_cfg):
    import svk_weights as W
    return W.dims_from_model_kwargs(513, **base_cfg["model"])


@pytest.fixture(scope="session")
def base_sd(base_dims):
    import json
    import svk_weights as W
    sd = W.make_state_dict(base_dims, seed=1234)
    meta = json.load(open(os.path.join(GOLDEN, "meta.json")))
    assert W.state_dict_checksum(sd) == meta["iitp_base_seed1234_checksum"], "weight recipe drifted"
    return sd


@pytest.fixture(scope="session")
def tiny_dims():
    import svk_weights as W
    return W.dims_from_model_kwargs(513, **W.TINY_MODEL)


@pytest.fixture(scope="session")
def tiny_sd(tiny_dims):
    import json
    import svk_weights as W
    sd = W.make_state_dict(tiny_dims, seed=4321)
    meta = json.load(open(os.path.join(GOLDEN, "meta.json")))
    assert W.state_dict_checksum(sd) == meta["tiny_seed4321_checksum"], "weight recipe drifted"
    return sd

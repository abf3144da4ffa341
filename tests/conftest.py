import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = "markov-process-analysis-on-point-cloud_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_ops():
    return np.load(os.path.join(GOLDEN, "ops.npz"))


@pytest.fixture(scope="session")
def golden_blocks():
    return np.load(os.path.join(GOLDEN, "blocks.npz"))


@pytest.fixture(scope="session")
def golden_models():
    return np.load(os.path.join(GOLDEN, "models.npz"))


@pytest.fixture(scope="session")
def golden_specs():
    with open(os.path.join(GOLDEN, "specs.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    from oracle import markov_oracle

    markov_oracle.build()
    return markov_oracle


@pytest.fixture(scope="session")
def mpc():
    """The product package (its directory name is not a Python identifier)."""
    return importlib.import_module(PKG)


def tape_of(npz, prefix):
    keys = sorted(k for k in npz.files if k.startswith(prefix + "_"))
    return [(k.rsplit("_", 1)[1], npz[k].astype(np.int64)) for k in keys]

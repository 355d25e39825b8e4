import json
import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long CPU oracle runs, opt in with PESTO_SLOW=1")


def _gpu_unavailable():
    if not torch.cuda.is_available():
        return "no CUDA device"
    from pesto_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        return f"{_lib.LIB_PATH} has not been built"
    return None


def pytest_collection_modifyitems(config, items):
    why = _gpu_unavailable()
    skip_gpu = pytest.mark.skip(reason=f"gpu test: {why}")
    skip_slow = pytest.mark.skip(reason="slow oracle run; set PESTO_SLOW=1")
    for item in items:
        if why and "gpu" in item.keywords:
            item.add_marker(skip_gpu)
        if "slow" in item.keywords and os.environ.get("PESTO_SLOW") != "1":
            item.add_marker(skip_slow)


def load_case(name):
    c = dict(np.load(os.path.join(GOLDEN, f"case_{name}.npz")))
    if "ids0" not in c and "ids0_0" in c and len(c["sizes"]) == 1:
        c["ids0"] = c["ids0_0"]          # single-structure cases written through the multi-structure path
    return c


def load_weights(tag):
    return dict(np.load(os.path.join(GOLDEN, f"weights_{tag}.npz")))


def load_config(tag):
    with open(os.path.join(GOLDEN, f"config_{tag}.json")) as fh:
        return json.load(fh)


def case_tensors(c):
    """(X, el, rid, n_res) torch tensors of a golden case."""
    return (torch.from_numpy(c["X"]), torch.from_numpy(c["el"].astype(np.int64)),
            torch.from_numpy(c["rid"].astype(np.int64)), int(c["n_res"]))


@pytest.fixture(scope="session")
def weights():
    cache = {}

    def get(tag):
        if tag not in cache:
            cache[tag] = load_weights(tag)
        return cache[tag]
    return get


@pytest.fixture(scope="session")
def cuda_models():
    """pesto_b200.Model instances on cuda:0 with the shipped checkpoints (from the golden weight fixtures).
    `mode` is the instance's default arithmetic: "fp32" (FFMA kernels) or "f16x3" (tcgen05, the mode bench.py times)."""
    from pesto_b200.model import Model
    cache = {}

    def get(tag, mode="fp32"):
        if (tag, mode) not in cache:
            sd = {k: torch.from_numpy(v) for k, v in load_weights(tag).items()}
            m = Model.for_state_dict(load_config(tag), sd, mode=mode)        # head depths read off the checkpoint (i_v3_1: one Linear)
            cache[(tag, mode)] = m.eval().to("cuda")
        return cache[(tag, mode)]
    return get


# the two arithmetic modes every size / property test runs in: the FFMA reference mode and the timed tensor-core mode
BOTH_MODES = ["fp32", "f16x3"]

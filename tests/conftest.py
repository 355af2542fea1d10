import ast
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """Golden fixture written by tests/golden/make_golden.py (outputs of the reference itself)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["meta"] = ast.literal_eval(str(d["meta"]))
    return d


def config_from_meta(meta, batch=1, device="cpu"):
    import amp_sparc_spatialmodulation_b200 as pkg
    Nt, Na, Nr, Lin, Lh, alphabet = meta["args"]
    kw = meta["kwargs"]
    return pkg.Config(Nt, Na, Nr, Lin, Lh, batch=batch, generator_mode=kw.get("mode", "sparc"),
                      iterations=kw.get("iters", 20), alphabet=alphabet, channel_profile="uniform",
                      channel_truncation=kw.get("trunc", "trunc"), device=device)


@pytest.fixture(scope="session")
def golden():
    return load_golden

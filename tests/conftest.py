import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle, with its C restatement built on demand."""
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    from oracle import oracle as o
    return o


def golden_inputs(g, name):
    """Inputs of a golden scan: stored, or regenerated from the seeded generator."""
    from powerfit_b200 import synth
    if "target" in g.files:
        return (g["target"].astype(np.float64), g["template"].astype(np.float64),
                g["mask"].astype(np.float64))
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    if "case_kwargs" in g.files:
        import json
        kw = json.loads(str(g["case_kwargs"]))
        if "shape" in kw:
            kw["shape"] = tuple(kw["shape"])
        case = synth.make_case(**kw)
        return f32(case.target), f32(case.template), f32(case.mask)
    cw = "cw" in name
    case = synth.config2(seed=int(g["seed"]), core_weighted=cw)
    return f32(case.target), f32(case.template), f32(case.mask)

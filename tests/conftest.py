import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name="cloze_tiny.npz"):
    z = np.load(os.path.join(ROOT, "tests", "golden", name))
    g = {"config": json.loads(bytes(z["config_json"]).decode()), "rows": int(z["rows"]),
         "randval": z["randval"], "task_weights": [float(x) for x in z["task_weights"]],
         "losses": z["losses"], "losses_bf16": z["losses_bf16"], "eval_flat": z["eval_flat"],
         "batch": {}, "masked": {}, "param": {}, "grad": {}, "grad_bf16": {}}
    for k in z.files:
        if "/" in k:
            grp, key = k.split("/", 1)
            g[grp][key] = z[k]
    return g


@pytest.fixture(scope="session")
def golden():
    return load_golden()

import json
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
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def example_fit():
    d = load_golden("example_fit.json")
    i = d["intr"]
    d["intr_tuple"] = (i["frow"], i["fcol"], i["crow"], i["ccol"], i["k"], i["checker_size"])
    d["view_list"] = [(v["rvec"], v["tvec"]) for v in d["views"]]
    d["corners_np"] = np.asarray(d["corners"], dtype=np.float64)  # (nviews, n1*n2, 2)
    d["obj_np"] = np.asarray(d["obj"], dtype=np.float64)
    return d


# synthetic cameras of SURVEY.md section 8(d)
C2_INTR = (1400.0, 1400.0, 540.0, 960.0, -0.12, 1.0)
C3_INTR = (2800.0, 2800.0, 1080.0, 1920.0, -0.12, 1.0)
SYN_VIEW = ((0.15, -0.1, 0.02), (-8.0, -12.0, 30.0))      # 85 % of output pixels in bounds
# bench view: board centred and nearly frontal so >= 99 % of output pixels sample in bounds
# (otherwise fewer bytes are read than the roofline counts)
BENCH_VIEW = ((0.05, -0.04, 0.02), (-9.3, -6.4, 30.0))


def camera_for(sz, k=-0.12, cs=1.0):
    """The C2/C3 camera family scaled to a frame of size sz: f = (1400/1080) sz1, principal
    point at the frame centre (C2_INTR == camera_for((1080, 1920)), C3_INTR == camera_for((2160, 3840)))."""
    f = 1400.0 / 1080.0 * sz[0]
    return (f, f, sz[0] / 2.0, sz[1] / 2.0, k, cs)

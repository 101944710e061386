import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE_ROOT = os.environ.get("JRR_REFERENCE_ROOT", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def jrr():
    import jrr_b200
    return jrr_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import jrr_oracle
    return jrr_oracle


@pytest.fixture(scope="session")
def model(jrr):
    return jrr.synthetic.make_smpl_model(0)


@pytest.fixture(scope="session")
def osmpl32(oracle, model):
    return oracle.OracleSMPL(model, torch.float32)


@pytest.fixture(scope="session")
def osmpl64(oracle, model):
    return oracle.OracleSMPL(model, torch.float64)


def shipped_regressor():
    """models/retrained_J_Regressor.pt rebuilt from its 107 non-zeros (tests/golden fixture
    written by make_golden.py from the reference artefact)."""
    z = np.load(os.path.join(GOLDEN, "j_regressor_nnz.npz"))
    J = np.zeros((17, 6890), dtype=np.float32)
    J[z["row"], z["col"]] = z["val"]
    return torch.from_numpy(J)


@pytest.fixture(scope="session")
def J_shipped():
    return shipped_regressor()


@pytest.fixture(scope="session")
def J_dense(jrr):
    return torch.from_numpy(jrr.synthetic.make_dense_regressor(0))


@pytest.fixture(scope="session")
def critic_sd(oracle):
    return oracle.make_critic_state_dict(0)


def make_frames(jrr, oracle, osmpl, J, n, seed):
    inp = jrr.synthetic.make_pose_inputs(n, seed)
    t = {k: torch.from_numpy(v) for k, v in inp.items()}
    t["gt_mm"] = oracle.make_gt(osmpl, J, t["true_rotmat"], t["true_betas"], t["gt_noise"])
    return t


@pytest.fixture(scope="session")
def frames64(jrr, oracle, osmpl32, J_shipped):
    return make_frames(jrr, oracle, osmpl32, J_shipped, 64, 0)

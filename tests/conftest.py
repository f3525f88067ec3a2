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
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


GOLDEN_CASES = {
    "cfg1_d7": dict(num_dof=7, num_basis=10, seq_len=50, vocab_size=256, degree_p=4,
                    gripper_zero_order=False, gripper_indices=[6], llm_vocab_size=None),
    "cfg2_d14": dict(num_dof=14, num_basis=10, seq_len=50, vocab_size=256, degree_p=4,
                     gripper_zero_order=True, gripper_indices=[6, 13], llm_vocab_size=32000),
    "odd_d5": dict(num_dof=5, num_basis=8, seq_len=33, vocab_size=1000, degree_p=3,
                   gripper_zero_order=True, gripper_indices=[0], llm_vocab_size=None),
    "cli_default": dict(num_dof=32, num_basis=50, seq_len=10, vocab_size=1000, degree_p=0,
                        gripper_zero_order=False, gripper_indices=None, llm_vocab_size=None),
}


@pytest.fixture(params=sorted(GOLDEN_CASES))
def golden_case(request):
    return request.param, GOLDEN_CASES[request.param], load_golden(request.param)


def rel_err(a, b):
    """max |a-b| / max |b| — the 1e-5 'relative tolerance' of the north star is
    read normwise (per tensor), as in SURVEY.md §0 trap 2."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

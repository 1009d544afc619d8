import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
WEIGHTS_DIR = os.path.join(ROOT, 'dmpfold2_b200', 'trained_model')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def have_trained_weights():
    return os.path.isfile(os.path.join(WEIGHTS_DIR, 'FINAL_fullmap_e2e_model_part1.pt')) and \
        os.path.isfile(os.path.join(WEIGHTS_DIR, 'FINAL_fullmap_e2e_model_part2.pt'))


needs_weights = pytest.mark.skipif(not have_trained_weights(), reason='trained weights not staged (run __graft_entry__.build())')


@pytest.fixture(scope='session')
def state_dict():
    from oracle import dmpfold_oracle as O
    if have_trained_weights():
        return O.load_state_dict(WEIGHTS_DIR)
    return O.random_state_dict(0)


@pytest.fixture(scope='session')
def oracle(state_dict):
    from oracle import dmpfold_oracle as O
    return O.Oracle(state_dict)


@pytest.fixture(scope='session')
def pf10963():
    from oracle import dmpfold_oracle as O
    return O.encode_aln(O.read_aln(os.path.join(GOLDEN, 'PF10963.aln')))

import os
import sys

import numpy as np
import pytest
import scipy.sparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def csr_from(npz, prefix):
    shape = tuple(int(v) for v in npz[prefix + '_shape'])
    return scipy.sparse.csr_matrix((npz[prefix + '_data'], npz[prefix + '_indices'], npz[prefix + '_indptr']),
                                   shape=shape)


def same_csr(A, B):
    A = scipy.sparse.csr_matrix(A)
    B = scipy.sparse.csr_matrix(B)
    A.sort_indices()
    B.sort_indices()
    return (A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
            and np.array_equal(A.data, B.data) and A.dtype == B.dtype)


@pytest.fixture(scope='session')
def c2():
    return load_golden('c2_grid28.npz')


@pytest.fixture(scope='session')
def c1():
    return load_golden('c1_usage.npz')


@pytest.fixture(scope='session')
def directed():
    return load_golden('directed57.npz')

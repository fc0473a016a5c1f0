import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import sbref
    sbref.build()
    return sbref


@pytest.fixture(scope="session")
def obg5(oracle):
    """Oracle background for the reference's test fixture: ΛCDM(lmax = 5), Planck18 (test/runtests.jl:14-17)."""
    return oracle.Background(oracle.planck18(lmax=5))


@pytest.fixture(scope="session")
def sb():
    import symboltz.jl_b200 as sb
    return sb


@pytest.fixture(scope="session")
def prob5(sb):
    M = sb.ΛCDM(lmax=5)
    return sb.CosmologyProblem(M, sb.parameters_Planck18(M))


@pytest.fixture(scope="session")
def bg5(sb, prob5):
    return sb.solvebg(prob5)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

INP = os.path.join(ROOT, "tests", "golden", "inp")
NET_A = os.path.join(INP, "rate06_dipole_reformated_again_withgrain.dat")
NET_B = os.path.join(INP, "rate12_withGrain_lowH2Bind_hiObind.dat")
NET_C = os.path.join(INP, "rate06_withgrain_lowH2Bind_hiOBind_lowCObind.dat")
IC_GARROD = os.path.join(INP, "initial_condition_Garrod08_mod_waterice.dat")
IC_LOMETAL = os.path.join(INP, "ini_abund_waterice_loMetal_CO.dat")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import raco
    raco.build()
    return raco


@pytest.fixture(scope="session")
def rb():
    import rac2d_b200
    rac2d_b200.build()
    return rac2d_b200

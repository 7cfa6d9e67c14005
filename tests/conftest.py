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
    return np.load(os.path.join(GOLDEN, name))


class Cosmo:
    """Host inputs of one cosmology (Background + RECFAST ionization history) and their packed form."""

    def __init__(self, **kw):
        import bolt_b200 as B
        import hostgen as HG
        from bolt_b200 import abi
        self.par = B.CosmoParams(**kw)
        self.bg = HG.Background(self.par)
        self.rec = HG.RECFAST(self.bg, OmegaB=self.par.Ω_b, Yp=self.par.Y_p, OmegaG=self.par.Ω_r)
        self.ih = HG.IonizationHistory(self.rec, self.par, self.bg)
        self.hc = abi.HostCosmo.from_host(self.par, self.bg, self.ih)
        self.ix_start = int(np.argmax(self.bg.x_grid > -8))


@pytest.fixture(scope="session")
def cosmo():
    """Default CosmoParams() (src/Bolt.jl:56-66), massive neutrinos included."""
    return Cosmo()


@pytest.fixture(scope="session")
def cosmo_nonu():
    """CosmoParams(Σm_ν=0) as in the reference's CLASS comparison (test/runtests.jl:85)."""
    return Cosmo(Σm_ν=0.0)


@pytest.fixture(scope="session")
def oracle(cosmo):
    from oracle.oracle import OracleCosmo
    return OracleCosmo(cosmo.hc)


@pytest.fixture(scope="session")
def gpu_ctx():
    from bolt_b200 import capi
    return capi.Context(0)


@pytest.fixture(scope="session")
def dev(cosmo, gpu_ctx):
    from bolt_b200 import capi
    return capi.DeviceCosmo(gpu_ctx, cosmo.hc)

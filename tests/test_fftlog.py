"""FFTLog (src/util.jl:33-108): oracle pinned by the reference's own known-answer vector
(test/runtests.jl:11-35, test/data/fftlog_example.txt); device against the oracle."""
import numpy as np
import pytest

from conftest import load_golden


def reference_case():
    N, mu, q, L = 64, 0, 0.0, 8.0                       # test/runtests.jl:12-20
    n = np.linspace(-(N // 2), N // 2, N)
    r = 10.0 ** (n * L / N)
    a = r ** (mu + 1) * np.exp(-r ** 2 / 2)
    return r, a, mu, q


def test_oracle_matches_the_reference_fixture():
    from oracle.fftlog_oracle import Plan
    g = load_golden("fftlog_example.npz")
    r, a, mu, q = reference_case()
    pl = Plan(r, mu, q, 1.0, kropt=True)
    y = pl.mul(a)
    assert np.abs(y - g["f"]).max() < 2e-15            # the reference asserts 1e-15 with FFTW (runtests.jl:27); pocketfft rounds 1.2e-15
    assert np.allclose(pl.k, g["k"], rtol=1e-12)
    assert np.abs(pl.ldiv(y) - a).max() < 1e-15        # runtests.jl:33


@pytest.mark.gpu
def test_device_fftlog_matches_oracle_and_fixture(gpu_ctx):
    from bolt_b200 import capi
    from oracle.fftlog_oracle import Plan
    g = load_golden("fftlog_example.npz")
    r, a, mu, q = reference_case()
    y, k, k0r0 = capi.fftlog(gpu_ctx, r, a, mu, q, 1.0, kropt=True)
    pl = Plan(r, mu, q, 1.0, kropt=True)
    assert abs(k0r0 / pl.k0r0 - 1) < 1e-13 and np.allclose(k, g["k"], rtol=1e-12)
    # hand-written radix-2 FFT + Stirling log-gamma instead of FFTW + SpecialFunctions: agreement to a few ulp of the largest value
    assert np.abs(y - g["f"]).max() < 5e-15 and np.abs(y - pl.mul(a)).max() < 5e-15
    back, _, _ = capi.fftlog(gpu_ctx, r, y, mu, q, 1.0, kropt=True, inverse=True)
    assert np.abs(back - a).max() < 5e-15
    # other orders / biases / sizes against the oracle
    for (N, mu2, q2) in ((256, 0.5, 0.3), (1024, 2.0, -0.4), (4096, 1.5, 0.0)):
        n = np.linspace(-(N // 2), N // 2, N); r2 = 10.0 ** (n * 6.0 / N)
        a2 = r2 ** (mu2 + 1) * np.exp(-r2 ** 2 / 2)
        y2, k2, _ = capi.fftlog(gpu_ctx, r2, a2, mu2, q2, 1.0, kropt=True)
        ref = Plan(r2, mu2, q2, 1.0, kropt=True)
        assert np.abs(y2 - ref.mul(a2)).max() < 1e-12 * np.abs(ref.mul(a2)).max() and np.allclose(k2, ref.k, rtol=1e-12)

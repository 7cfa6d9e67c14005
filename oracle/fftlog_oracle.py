"""CPU oracle of the FFTLog row (TEST INFRASTRUCTURE ONLY): a numpy/scipy restatement of src/util.jl:33-108, function by function.
Pinned by the reference's own fixture test/data/fftlog_example.txt through test/runtests.jl:11-35 (tests/test_fftlog.py)."""
import numpy as np
from scipy.special import loggamma


def U_mu(mu, x):                                             # util.jl:76
    return np.exp(x * np.log(2.0) - loggamma(0.5 * (mu + 1 - x)) + loggamma(0.5 * (mu + 1 + x)))


def u_m(m, mu, q, dlnr, k0r0, N):                            # util.jl:77
    return k0r0 ** (-2j * np.pi * m / (dlnr * N)) * U_mu(mu, q + 2j * np.pi * m / (dlnr * N))


def k0r0_low_ringing(N, mu, q, L, k0r0=1.0):                 # util.jl:79-89
    dlnr = L / (N - 1)
    xp, xm = (mu + 1 + q) / 2, (mu + 1 - q) / 2
    y = np.pi * 1j / 2 / dlnr
    zp, zm = loggamma(xp + y), loggamma(xm + y)
    arg = np.log(2 / k0r0) / dlnr + (zp + zm).imag / np.pi
    return k0r0 * np.exp((arg - np.round(arg)) * dlnr)


class Plan:                                                  # plan_fftlog, util.jl:45-72
    def __init__(self, r, mu, q, k0r0=1.0, kropt=True):
        r = np.asarray(r, dtype=np.float64)
        logrmin, logrmax = np.log(r[0]), np.log(r[-1])
        r0 = np.exp((logrmin + logrmax) / 2)
        N = len(r); L = logrmax - logrmin; dlnr = L / (N - 1)
        if kropt:
            k0r0 = k0r0_low_ringing(N, mu, q, L, k0r0)
        k0 = k0r0 / r0
        n = np.linspace(-(N // 2), N // 2, N)
        self.k = (k0 * np.exp(n * L / N))[::-1]
        m = np.fft.fftfreq(N, 1.0 / N)
        um = u_m(m, mu, q, dlnr, k0r0, N).astype(np.complex128)
        um[N // 2] = um[N // 2].real                          # eq. 19
        self.r, self.q, self.um, self.k0r0, self.N = r, q, um, k0r0, N

    def mul(self, a):                                        # mul!, util.jl:91-98
        y = np.asarray(a, dtype=np.complex128) * self.r ** (-self.q)
        return np.fft.ifft(np.fft.fft(y) * self.um) * self.r ** self.q

    def ldiv(self, a):                                       # ldiv!, util.jl:100-107
        y = np.asarray(a, dtype=np.complex128) * self.r ** (-self.q)
        return np.fft.ifft(np.fft.fft(y) / self.um) * self.r ** self.q

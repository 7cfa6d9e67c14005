#!/usr/bin/env python
"""Benchmark of the hot path on B200 (contract: one JSON line on stdout from rank 0).

Workload (config.workload = "C3-value"): BASELINE.json configs[2] without gradients -- per cosmology 2000
quadratic k-modes (0.1 H0 .. 1000 H0), l_gamma = 8 (state n = 197, massive neutrinos always in the state),
adaptive KenCarp4 at reltol 1e-11 / abstol 1e-6 (the reference's source_grid defaults), then TT/TE/EE C_l
for l = 2..2500 on the 5000-point dense k grid.  A "step" = that whole pipeline for ONE synthetic cosmology
per GPU.  Multi-GPU: cosmologies shard over ranks with no data-path collective (weak scaling).

  value  = k-mode hierarchy solves/s of the whole job (all ranks), spline tables already resident in HBM
  e2e    = same metric through the reference-facing API with HOST buffers: every step uploads the cosmology's
           tables (H2D), runs bolt_spectra and reads C_l back (D2H)
  --impl reference : the oracle port (C++/OpenMP, all host cores) on a bounded sample of the same workload
  --scaling strong : BASELINE configs[3] (10^4 modes, l_gamma = 50, ONE cosmology) sharded over the GPUs inside the library

Extra objects on the line (never at the cost of the headline: each arm is guarded): `gradients` (the C3 workload with six forward-mode
partials and its own CPU baseline -- the workload north_star's >= 100x target is quoted on), `plin` (configs[1]), `batch`
(bolt_spectra_batch), `hostgen` and `params_to_spectra` (input tables of a batch of parameter sets made on the device), `filon`
(Bessel-moment table + Filon chains, with the numpy oracle as its CPU arm); at N > 1 `strong_scaling` (ONE cosmology sharded inside the
library with NCCL: bolt_spectra_sharded, bolt_plin_sharded).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NK, ELL_MIN, ELL_MAX, LG, RELTOL, ABSTOL = 2000, 2, 2500, 8, 1e-11, 1e-6
SEED = 20261017
GRAD_NAMES = ["Ω_b", "Ω_c", "h", "n", "A", "Σm_ν"]


def synthetic_params(i):
    """Synthetic ΛCDM(+massive ν) parameter draws (SURVEY 8d ranges), deterministic in (SEED, i)."""
    import bolt_b200 as B
    import hostgen as HG
    from hostgen import constants as K
    rng = np.random.default_rng([SEED, i])
    u = rng.random(6)
    return B.CosmoParams(h=0.60 + 0.20 * u[0], Ω_b=0.040 + 0.015 * u[1], Ω_c=0.20 + 0.10 * u[2], n=0.92 + 0.08 * u[3],
                         A=1e-10 * np.exp(2.9 + 0.3 * u[4]), Σm_ν=0.3 * u[5] * K.mass_natural)


def make_host_cosmo(i):
    import bolt_b200 as B
    import hostgen as HG
    from bolt_b200 import abi
    par = synthetic_params(i)
    bg = HG.Background(par)
    ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
    hc = abi.HostCosmo.from_host(par, bg, ih)
    k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, NK)
    ix_start = int(np.argmax(bg.x_grid > -8))
    return dict(par=par, bg=bg, hc=hc, k=k, ix_start=ix_start, kd=(0.01 * bg.H0, 1000 * bg.H0, 5000))


def f_step(n):
    """Algorithmic FP64 flop of one (accepted or rejected) KenCarp4 step of the structured solver:
    5 stages x (assemble 6n + factor 6n + solve 8n + stage increment 8n + border 600) + error pass 22n +
    smoothing solve 14n + 300 + norm 6n."""
    return 182.0 * n + 3300.0


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 10 ms from a thread of this process
    (nvidia_ml_py); `nvidia-smi -lms` in a child process only if NVML cannot be loaded (its first sample takes ~0.5 s, longer
    than a short timed region)."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            index = int(vis.split(",")[index]) if vis else index
        except (ValueError, IndexError):
            pass
        self.index, self.rows, self.proc, self.nv, self.stop_flag = index, [], None, None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True); self.t.start()
            return
        except Exception:
            self.nv = None
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = [0x8, 0x40, 0x20, 0x4]          # HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap (nvml.h)
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)); r = int(get_reasons(self.h))
                self.rows.append([str(sm), str(self.mx)] + ["Active" if r & b else "Not Active" for b in bits])
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["neither NVML nor nvidia-smi available"], "samples": 0}
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = [nm for j, nm in enumerate(self.NAMES) if any(len(r) >= 6 and r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nv is not None else "nvidia-smi"}


def host_threads():
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its children: the CPU arm overrides it
    (oracle_set_num_threads), otherwise the OpenMP oracle would run single-threaded at N > 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(hcosmo, kstride=50):
    """One bounded sample of the step on the oracle port (all host cores): every kstride-th k-mode through the adaptive
    solve and the SAME fraction of the multipoles through the projection.  Returns k-mode solves/s of the sample (the
    sample's solves / its wall time: no extrapolation to the full step is involved) and the sample's wall time."""
    from bolt_b200 import abi
    from oracle.oracle import OracleCosmo, lib
    lib().oracle_set_num_threads(host_threads())
    oc = OracleCosmo(hcosmo["hc"])
    cores = lib().oracle_num_threads()
    o = abi.make_opts(LG, 8, 10, reltol=RELTOL, abstol=ABSTOL, ix_first=hcosmo["ix_start"])
    ks = np.ascontiguousarray(hcosmo["k"][kstride // 2::kstride])
    t0 = time.perf_counter(); out = oc.solve(ks, o, want=("S_T", "S_P")); t_solve = time.perf_counter() - t0
    nell_full = ELL_MAX - ELL_MIN + 1
    ells = np.unique(np.linspace(ELL_MIN, ELL_MAX, max(2, round(nell_full * len(ks) / NK))).astype(np.int32))
    kmin, kmax, nkd = hcosmo["kd"]
    t0 = time.perf_counter(); oc.project(out["S_T"], out["S_P"], ks, ells, kmin, kmax, nkd, hcosmo["ix_start"]); t_proj = time.perf_counter() - t0
    t = t_solve + t_proj
    return dict(value=len(ks) / t, unit="k-mode solves/s", cores=cores, kind="port",
                sample=f"every {kstride}th of the {NK} k-modes ({len(ks)} adaptive solves, {t_solve:.1f}s) + the same fraction of the multipoles "
                       f"({len(ells)} of {nell_full}, {t_proj:.1f}s), OpenMP over k and l on {cores} threads; value = sample solves / sample time; "
                       "C++ restatement with zero-skipping dense LU (faster than the reference's dense LU), not Julia"), t


def cpu_sample_gradients(dual, bg, kstride=125):
    """The gradient workload (C3 with six forward-mode partials) on the oracle's dual-number stepper (generic dual arithmetic
    on the restated right-hand side + dense LU, what the reference does with ForwardDiff.Dual parameters), bounded sample."""
    from bolt_b200 import abi
    from oracle.oracle import OracleCosmo, lib
    import bolt_b200 as B
    import hostgen as HG
    lib().oracle_set_num_threads(host_threads())
    oc = OracleCosmo(dual)
    cores = lib().oracle_num_threads()
    k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, NK)
    ix0 = int(np.argmax(bg.x_grid > -8))
    o = abi.make_opts(LG, 8, 10, reltol=RELTOL, abstol=ABSTOL, ix_first=ix0)
    ks = np.ascontiguousarray(k[kstride // 2::kstride])
    t0 = time.perf_counter(); out = oc.solve_sens(ks, o, want=("S_T", "S_P")); t_solve = time.perf_counter() - t0
    nell_full = ELL_MAX - ELL_MIN + 1
    ells = np.unique(np.linspace(ELL_MIN, ELL_MAX, max(2, round(nell_full * len(ks) / NK))).astype(np.int32))
    t0 = time.perf_counter(); oc.project_sens(out["S_T"], out["S_P"], ks, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0); t_proj = time.perf_counter() - t0
    t = t_solve + t_proj
    return dict(value=len(ks) / t, unit="k-mode solves/s (value + 6 partials each)", cores=cores, kind="port",
                sample=f"every {kstride}th of the {NK} k-modes ({len(ks)} dual-number solves, {t_solve:.1f}s) + the same fraction of the "
                       f"multipoles ({len(ells)} of {nell_full}, {t_proj:.1f}s), {cores} threads; value = sample solves / sample time"), t


def gradient_arm(ctx, ells, with_cpu=True):
    """BASELINE configs[2] proper and the workload north_star's >= 100x target is quoted on: the C3 workload with forward-mode
    partials w.r.t. six parameters (value + gradient from one pass, nd = 7).  tau is not a parameter of the reference:
    the sixth direction is the neutrino mass.  One warm-up and three timed steps (host buffers: the e2e form), and the oracle's
    dual-number stepper on a bounded sample of the same workload as its CPU baseline."""
    import bolt_b200 as B
    import hostgen as HG
    from bolt_b200 import abi, capi
    from hostgen import host_cosmo_with_partials
    par = synthetic_params(0)
    dual, base, bg, ih, pm, steps = host_cosmo_with_partials(par, GRAD_NAMES, rel_step=1e-3)
    dc = capi.DeviceCosmo(ctx, dual)
    k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, NK)
    o = abi.make_opts(LG, 8, 10, reltol=RELTOL, abstol=ABSTOL)
    ix0 = int(np.argmax(bg.x_grid > -8))
    out = dc.spectra(k, o, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    nrep = 3
    t0 = time.perf_counter()
    for rep in range(nrep):
        out = dc.spectra(k, o, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0)
    dt = (time.perf_counter() - t0) / nrep
    tm = ctx.timing()
    res = {"workload": "C3 with gradients: the C3 workload with 6 forward-mode partials (nd = 7), TT+TE+EE and their gradients",
           "params": GRAD_NAMES, "nd": 7, "ms_per_step": 1e3 * dt, "spectra_with_gradients_per_s": 1.0 / dt, "value": NK / dt,
           "unit": "k-mode solves/s (value + 6 partials each)", "timed_steps": nrep,
           "kernel_ms": {"hierarchy": tm["hierarchy_ms"], "bessel_tables": tm["bessel_ms"], "projection": tm["project_ms"]},
           "ode_steps_per_solve": float(out[4].mean()), "failed_modes": int((out[3] != 0).sum()),
           "note": "host tables' partials by central differences of the Python host generator (the Julia shim passes ForwardDiff "
                   "partials); error control runs over value and partials like the reference, hence more steps than value-only"}
    if with_cpu:
        cb, _ = cpu_sample_gradients(dual, bg)
        res["cpu_baseline"] = cb
        res["speedup_vs_cpu_baseline"] = res["value"] / cb["value"]
        res["north_star_target"] = ">= 100x the all-core CPU spectra/s for TT+EE with gradients at 1 B200"
    return res


def plin_arm(ctx, hcosmo, dc):
    """BASELINE configs[1]: linear matter P(k) via plin with massive neutrinos, 500 log-spaced k-modes, the reference's plin
    defaults l_gamma = l_nu = 50, l_mnu = 20 (state n = 473), reltol 1e-5 (src/spectra.jl:163-164).  Runtime-truncation K1 path."""
    import bolt_b200 as B
    import hostgen as HG
    from bolt_b200 import abi
    bg = hcosmo["bg"]
    ks = B.log10_k(10 * bg.H0, 5000 * bg.H0, 500)
    o = abi.make_opts(50, 50, 20, reltol=1e-5, abstol=1e-6)
    for rep in range(2):
        t0 = time.perf_counter(); pk, st, ns = dc.plin(ks, o); dt = time.perf_counter() - t0
    return {"workload": "plin, 500 log10_k modes (10 H0 .. 5000 H0), n = 473, reltol 1e-5", "ms": 1e3 * dt, "kmode_solves_per_s": len(ks) / dt,
            "hierarchy_ms": ctx.timing()["hierarchy_ms"], "ode_steps_per_solve": float(ns.mean()), "failed_modes": int((st != 0).sum())}


def params_to_spectra_arm(ctx, local_rank, ells, ncos=16):
    """BASELINE configs[4] in miniature, end to end ON THE DEVICE: parameter sets in, TT/TE/EE out -- input tables by
    bolt_hostgen_batch (one call for the batch), upload, bolt_spectra per cosmology.  What an emulator / MCMC driver would run."""
    import bolt_b200 as B
    from bolt_b200 import abi, capi
    pars = [synthetic_params(100 + i) for i in range(ncos)]
    o = abi.make_opts(LG, 8, 10, reltol=RELTOL, abstol=ABSTOL)
    t0 = time.perf_counter()
    hcs, st = capi.hostgen_batch(pars, device=local_rank)
    t_gen = time.perf_counter() - t0
    bad = 0
    for hc in hcs:
        H0 = hc.scalar("H0")
        dc = capi.DeviceCosmo(ctx, hc)
        k = B.quadratic_k(0.1 * H0, 1000 * H0, NK)
        out = dc.spectra(k, o, ells, 0.01 * H0, 1000 * H0, 5000, 1201)
        bad += int((out[3] != 0).sum()); dc.close()
    dt = time.perf_counter() - t0
    return {"workload": f"{ncos} parameter sets -> input tables on the device (one bolt_hostgen_batch call) -> upload -> C3-value spectra per cosmology",
            "ms_per_cosmology": 1e3 * dt / ncos, "spectra_per_s": ncos / dt, "table_generation_ms_per_call": 1e3 * t_gen,
            "failed_modes": bad, "failed_tables": int((st != 0).sum())}


def hostgen_arm(local_rank, ncos=256):
    """the input tables (background, RECFAST, reionization, optical depth and their spline coefficients) of a BATCH of
    synthetic cosmologies on the device (bolt_hostgen_batch: one thread integrates the recombination ODEs of one cosmology), beside
    the Python harness generator on one host core (the reference computes them on the host too, one cosmology at a time)."""
    import hostgen as HG
    from bolt_b200 import capi
    pars = [synthetic_params(i) for i in range(ncos)]
    capi.hostgen_batch(pars[:2], device=local_rank)
    t0 = time.perf_counter(); hcs, st = capi.hostgen_batch(pars, device=local_rank); dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    bg = HG.Background(pars[0]); HG.IonizationHistory(HG.RECFAST(bg, OmegaB=pars[0].Ω_b, Yp=pars[0].Y_p, OmegaG=pars[0].Ω_r), pars[0], bg)
    t_cpu = time.perf_counter() - t0
    return {"workload": f"input tables of {ncos} synthetic cosmologies in one call (12 spline tables x 2003 coefficients + 13 scalars each, "
                        "returned to host buffers)", "ms_per_call": 1e3 * dt, "ms_per_cosmology": 1e3 * dt / ncos, "cosmologies_per_s": ncos / dt,
            "failed": int((st != 0).sum()),
            "cpu_baseline": {"value": 1.0 / t_cpu, "unit": "cosmologies/s", "cores": 1, "kind": "port",
                             "sample": f"one cosmology through the Python harness generator hostgen/ ({t_cpu:.2f} s)"},
            "speedup_vs_cpu_baseline": (ncos / dt) * t_cpu}


def cpu_filon_chain(nodes, f, f1, f2, k, N, xmax):
    """CPU leg of filon_arm: the oracle builds its own ν = 2 moment table and runs the same chains (numpy, one core)."""
    from oracle import bessel_moments_oracle as BO
    t0 = time.perf_counter(); ora = BO.MomentTable(2, 3, 0.0, xmax, N); t_build = time.perf_counter() - t0
    t0 = time.perf_counter(); ref = BO.filon_chain(nodes, f, f1, f2, k, ora); t_run = time.perf_counter() - t0
    return ref, t_build, t_run


def filon_arm(local_rank, n_k=2000, n_nodes=2048, cpu_k=2000):
    """the reference's Bessel-moment table (src/bessel/interpolator.jl:67-110: 2,000,000 nodes on kη in (0, 1.6e4),
    ν = 2) built on the device, then a line-of-sight-shaped Filon workload: for each of n_k wavenumbers, the chain of quadratic
    source pieces between n_nodes conformal-distance nodes (integrator.jl:25-38).  CPU arm: the numpy-vectorised oracle doing
    the same gathers from a table it built itself (one core)."""
    import bolt_b200.bessel as BM
    N, XMAX = 2_000_000, 1.6e4
    BM.sph_bessel_interpolator(2, 3, 0.0, 100.0, 1000).close()                  # load + first-launch cost out of the way
    t0 = time.perf_counter(); itp = BM.sph_bessel_interpolator(2, 3, 0.0, XMAX, N, device=local_rank); t_build = time.perf_counter() - t0
    nodes = np.linspace(0.0, 14000.0, n_nodes)
    k = np.linspace(1e-4, 1.0, n_k)
    g = np.exp(-0.5 * ((nodes - 13800.0) / 60.0) ** 2)[None, :] * np.cos(k[:, None] * 0.3)
    d1 = -(nodes - 13800.0)[None, :] / 3600.0 * g
    d2 = (((nodes - 13800.0) ** 2 / 3600.0 - 1.0) / 3600.0)[None, :] * g
    out, ms = BM.filon_chain(nodes, g, d1, d2, k, itp, timing=True)
    t0 = time.perf_counter(); out, ms = BM.filon_chain(nodes, g, d1, d2, k, itp, timing=True); t_call = time.perf_counter() - t0
    itp.close()
    sel = np.linspace(0, n_k - 1, cpu_k).astype(int)
    ref, t_build_cpu, t_cpu = cpu_filon_chain(nodes, g[sel], d1[sel], d2[sel], k[sel], N, XMAX)
    # the chain sums cancel heavily (oscillatory integrand; the rule's c0, c1, c2 are powers of x, not of x - a): compare on the
    # scale of the largest chain
    err = float(np.max(np.abs(out[sel] - ref)) / np.abs(ref).max())
    pieces = n_k * (n_nodes - 1)
    return {"workload": f"moment table ν=2, order 3, {N} nodes on (0, {XMAX:g}); Filon chains: {n_k} wavenumbers x {n_nodes - 1} quadratic pieces",
            "table_build_ms": 1e3 * t_build, "chain_kernel_ms": ms, "chain_call_ms_host_buffers": 1e3 * t_call,
            "pieces_per_s_kernel": pieces / (1e-3 * ms), "table_gather_GBps": pieces * 128 / (1e-3 * ms) / 1e9,
            "max_diff_vs_oracle_over_largest_chain": err,
            "cpu_baseline": {"table_build_s": t_build_cpu, "value": cpu_k * (n_nodes - 1) / t_cpu, "unit": "pieces/s", "cores": 1, "kind": "port",
                             "sample": f"{cpu_k} of the {n_k} wavenumbers through oracle/bessel_moments_oracle.py (numpy, {t_cpu:.2f} s); "
                                       f"its table build took {t_build_cpu:.1f} s (40-digit sums for the 6250 nodes below kη = 50)"},
            "speedup_vs_cpu_baseline_kernel": (pieces / (1e-3 * ms)) / (cpu_k * (n_nodes - 1) / t_cpu)}


def batch_arm(ctx, hcos, dcs, ells, ncos=4):
    """SURVEY 8d C5 (emulator / MCMC batches): the same C3-value workload through bolt_spectra_batch, ncos cosmologies per call --
    all ncos x 2000 hierarchy solves in ONE launch, so the tail of the persistent kernel is paid once per batch."""
    from bolt_b200 import abi, capi
    sel = [i % 2 for i in range(ncos)]
    ks = np.stack([hcos[j]["k"] for j in sel])
    kmin = np.array([hcos[j]["kd"][0] for j in sel]); kmax = np.array([hcos[j]["kd"][1] for j in sel])
    o = abi.make_opts(LG, 8, 10, reltol=RELTOL, abstol=ABSTOL)
    for rep in range(2):
        t0 = time.perf_counter()
        out = capi.spectra_batch(ctx, [dcs[j] for j in sel], ks, o, ells, kmin, kmax, hcos[0]["kd"][2], hcos[0]["ix_start"])
        dt = time.perf_counter() - t0
    return {"workload": f"C3-value x {ncos} cosmologies per call (bolt_spectra_batch)", "ms_per_call": 1e3 * dt, "ms_per_cosmology": 1e3 * dt / ncos,
            "kmode_solves_per_s": ncos * NK / dt, "spectra_per_s": ncos / dt, "hierarchy_ms": ctx.timing()["hierarchy_ms"],
            "failed_modes": int((out[3] != 0).sum())}


def strong_main(args, ctx, rank, world, local_rank, barrier):
    """BASELINE configs[3]: high-resolution k-sweep of ONE cosmology -- 10^4 quadratic k-modes, l_gamma = 50 (state n = 281, the
    runtime-truncation K1 path), adaptive 1e-11, full LOS projection l = 2..2500 -- sharded over the GPUs inside the library
    (bolt_spectra_sharded).  Strong scaling: the total work is fixed, value = 10^4 / time per spectrum set."""
    import torch
    import torch.distributed as dist
    from bolt_b200 import abi, capi
    import bolt_b200 as B
    import hostgen as HG
    NK4, LG4 = 10000, 50
    par = synthetic_params(0)
    bg = HG.Background(par)
    ih = HG.IonizationHistory(HG.RECFAST(bg, OmegaB=par.Ω_b, Yp=par.Y_p, OmegaG=par.Ω_r), par, bg)
    hc = abi.HostCosmo.from_host(par, bg, ih)
    k = B.quadratic_k(0.1 * bg.H0, 1000 * bg.H0, NK4)
    ix0 = int(np.argmax(bg.x_grid > -8))
    ells = np.arange(ELL_MIN, ELL_MAX + 1, dtype=np.int32)
    o = abi.make_opts(LG4, 8, 10, reltol=RELTOL, abstol=ABSTOL)
    n = abi.state_dim(LG4, 8, 10, 15)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)        # NCCL's version banner goes to stderr
        try:
            ctx.comm_init_torch()
        finally:
            os.dup2(saved, 1); os.close(saved)
    dc = capi.DeviceCosmo(ctx, hc)
    call = (lambda d: d.spectra_sharded(k, o, ells, 0.01 * bg.H0, 1000 * bg.H0, 5000, ix0))

    def step_e2e():
        d = capi.DeviceCosmo(ctx, hc)        # H2D tables
        out = call(d)                        # H2D k, ells; D2H C_l, status, step counts
        d.close()
        return out

    for i in range(args.warmup):
        call(dc)
    fp64_peak = ctx.fp64_peak_tflops()
    sampler = ClockSampler(local_rank); sampler.start()
    barrier(); t0 = time.perf_counter()
    k1_ms = k2_ms = dev_ms = 0.0; launches = 0; attempts = 0; bad = 0
    for i in range(args.steps):
        tt, te, ee, st, ns = call(dc)
        tm = ctx.timing(); k1_ms += tm["hierarchy_ms"]; k2_ms += tm["project_ms"]; dev_ms += tm["total_ms"]
        launches += tm["hierarchy_launches"] + tm["bessel_launches"] + tm["project_launches"]
        attempts += int(ns.sum() + dc.last_nreject.sum()); bad += int((st != 0).sum())
    barrier(); wall = time.perf_counter() - t0
    clocks = sampler.stop()
    step_e2e(); barrier(); t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e()
    barrier(); wall_e2e = time.perf_counter() - t0
    times = torch.tensor([wall, wall_e2e, k1_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    wall, wall_e2e, k1_max = [float(v) for v in times.tolist()]
    if rank == 0:
        nell = len(ells)
        fl = attempts / world * f_step(n)     # every rank reports the all-gathered counts of ALL modes: its own share is 1/N of them
        line = {"metric": "kmode_hierarchy_solves_per_s", "value": NK4 * args.steps / wall, "unit": "k-mode solves/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C4: ONE cosmology, 10^4 quadratic k-modes (0.1 H0 .. 1000 H0), l_gamma = 50, l_nu = 8, l_mnu = 10 (n = 281), adaptive "
                                       "KenCarp4 reltol 1e-11 / abstol 1e-6, TT+TE+EE C_l for l = 2..2500 on the 5000-point dense k grid",
                           "parallelism": f"k-modes (K1) and multipoles (K2) sharded x{world} inside the library: ncclAllGather of the source columns "
                                          "(320 MB), one ncclAllReduce of C_l (60 KB)",
                           "l2": "no explicit flush: 320 MB of source grids + 200 MB of Bessel tables per step > 126 MB L2"},
                "spectra_per_s": args.steps / wall, "kernel_ms_per_step_rank0": {"hierarchy": k1_ms / args.steps, "projection": k2_ms / args.steps,
                                                                                  "device_total": dev_ms / args.steps},
                "ode_step_attempts_per_solve": attempts / (NK4 * args.steps), "failed_modes": bad, "clocks": clocks, "gpu_launches": launches,
                "roofline": {"kernel": "hierarchy_kernel_t<TruncRT> (K1, runtime truncations)", "bound": "fp64",
                             "achieved": fl / (k1_max * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": fl / (k1_max * 1e-3) / 1e12 / fp64_peak if fp64_peak else None, "traffic": None,
                             "note": "algorithmic flop = this rank's step attempts x (182 n + 3300), n = 281, over the slowest rank's K1 time"},
                "e2e": {"value": NK4 * args.steps / wall_e2e, "unit": "k-mode solves/s",
                        "h2d_bytes_per_step": int(world * (hc.tables.nbytes + hc.scalars.nbytes + 2 * 15 * 8 + NK4 * 8 + nell * 4)),
                        "d2h_bytes_per_step": int(world * (3 * nell * 8 + NK4 * 4 + 2 * NK4 * 8))}}
        print(json.dumps(line))
    if world > 1:
        ctx.comm_free()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gradients", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): cosmology-sharded C3-value; strong: BASELINE configs[3] (10^4 k-modes, l_gamma = 50, ONE cosmology) "
                         "with its k-modes and multipoles sharded over the GPUs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "C3-value: per cosmology 2000 quadratic k-modes (n=197, l_gamma=8, l_nu=8, l_mnu=10, nq=15), adaptive KenCarp4 "
                          "reltol 1e-11/abstol 1e-6, TT+TE+EE C_l for l=2..2500 on the 5000-point dense k grid; gradients not included",
              "cosmologies_per_step_per_gpu": 1,
              "parallelism": (f"cosmology-sharded x{args.gpus}, no data-path collective: distinct synthetic cosmologies pulled from a shared work queue "
                              "(steps x N jobs per timed region)") if args.gpus > 1 else "one GPU, two synthetic cosmologies alternated",
              "l2": "no explicit flush: each step re-creates >330 MB of intermediates (Bessel tables 200 MB, dense source grids 64 MB, "
                    "source grids 64 MB) > 126 MB L2, and alternates between two cosmologies"}

    if args.impl == "reference":
        # Rank 0 alone runs the CPU arm (with every host core, whatever OMP_NUM_THREADS torchrun exported); other ranks exit 0.
        if rank != 0:
            return
        hcos = make_host_cosmo(0)
        vals, secs = [], []
        for i in range(args.warmup + args.steps):       # every step = one bounded sample (cpu_sample docstring), ~2-3 s each
            cb, t = cpu_sample(hcos)
            if i >= args.warmup:
                vals.append(cb["value"]); secs.append(t)
        v = float(len(vals) * len(hcos["k"][25::50]) / np.sum(secs))          # sample solves of the timed steps / their time
        cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": "kmode_hierarchy_solves_per_s", "value": v, "unit": "k-mode solves/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": config, "cpu_baseline": cb,
                          "step_note": "a step of this arm is a bounded sample of the workload (1/50 of the k-modes and of the multipoles); "
                                       "value = sample k-mode solves / sample time, so ms_per_step is the sample's time, not a full spectrum set",
                          "e2e": {"value": v, "unit": "k-mode solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from bolt_b200 import abi, capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = capi.Context(local_rank)
    if args.scaling == "strong":
        strong_main(args, ctx, rank, world, local_rank, barrier)
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    # N = 1: two synthetic cosmologies, alternated between steps.
    # N > 1: a pool of DISTINCT synthetic cosmologies behind a SHARED WORK QUEUE (an atomic counter in the rendezvous store):
    # the timed region is one batch of steps x N cosmologies (job j = pool[j % pool size]); a rank pulls the next job when it is
    # free, so the +-40 % cost spread between parameter draws (number of ODE steps) is balanced by the queue instead of being
    # hidden by giving every rank identical work.  No collective on the data path.
    npool = 2 if world == 1 else 4
    hcos = [make_host_cosmo(j) for j in range(npool)]
    dcs = [capi.DeviceCosmo(ctx, h["hc"]) for h in hcos]
    ells = np.arange(ELL_MIN, ELL_MAX + 1, dtype=np.int32)
    n = abi.state_dim(LG, 8, 10, 15)
    store = dist.distributed_c10d._get_default_store() if world > 1 else None

    def jobs(tag, njobs):
        """Job indices for this rank: the local loop at N = 1, pulls from the shared counter at N > 1."""
        if world == 1:
            yield from range(njobs)
            return
        while True:
            j = store.add(f"bolt_bench_{tag}", 1) - 1
            if j >= njobs:
                return
            yield j

    def step(j):
        h, dc = hcos[j % npool], dcs[j % npool]
        o = abi.make_opts(LG, 8, 10, reltol=RELTOL, abstol=ABSTOL)
        kmin, kmax, nkd = h["kd"]
        tt, te, ee, st, ns = dc.spectra(h["k"], o, ells, kmin, kmax, nkd, h["ix_start"])
        return tt, te, ee, st, ns + dc.last_nreject, ctx.timing()

    def step_e2e(j):
        h = hcos[j % npool]
        dc = capi.DeviceCosmo(ctx, h["hc"])            # H2D: tables + scalars + quadrature
        o = abi.make_opts(LG, 8, 10, reltol=RELTOL, abstol=ABSTOL)
        kmin, kmax, nkd = h["kd"]
        out = dc.spectra(h["k"], o, ells, kmin, kmax, nkd, h["ix_start"])   # H2D k, ells ; D2H C_l, status, nsteps
        dc.close()
        return out

    for i in range(args.warmup):
        step(i + rank)
    fp64_peak = ctx.fp64_peak_tflops()
    sampler = ClockSampler(local_rank); sampler.start()
    njobs = args.steps * world
    barrier()
    t0 = time.perf_counter()
    dev_ms, k1_ms, k2_ms, bes_ms, flops, launches, nstep_tot = 0.0, 0.0, 0.0, 0.0, 0.0, 0, 0
    bad = 0; my_jobs = 0
    for j in jobs("dev", njobs):
        tt, te, ee, st, ns, tm = step(j)
        dev_ms += tm["total_ms"]; k1_ms += tm["hierarchy_ms"]; k2_ms += tm["project_ms"]; bes_ms += tm["bessel_ms"]
        launches += tm["hierarchy_launches"] + tm["bessel_launches"] + tm["project_launches"]
        nstep_tot += int(ns.sum()); bad += int((st != 0).sum()); my_jobs += 1
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    # e2e arm
    for i in range(2):
        step_e2e(i + rank)
    barrier()
    t0 = time.perf_counter()
    for j in jobs("e2e", njobs):
        step_e2e(j)
    barrier()
    wall_e2e = time.perf_counter() - t0

    # Strong scaling of ONE cosmology (N > 1): the same C3-value workload with its k-modes (K1) and multipoles (K2) sharded over
    # the ranks inside the library -- ncclAllGather of the source columns, one ncclAllReduce of C_l (bolt_spectra_sharded)
    strong = None
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)        # NCCL prints its version banner on stdout: keep stdout to the ONE JSON line
        try:
            ctx.comm_init_torch()
        finally:
            os.dup2(saved, 1); os.close(saved)
        h, dc = hcos[0], dcs[0]
        o = abi.make_opts(LG, 8, 10, reltol=RELTOL, abstol=ABSTOL)
        kmin, kmax, nkd = h["kd"]
        outs = dc.spectra_sharded(h["k"], o, ells, kmin, kmax, nkd, h["ix_start"])
        barrier()
        t0 = time.perf_counter()
        nrep = 5
        for rep in range(nrep):
            outs = dc.spectra_sharded(h["k"], o, ells, kmin, kmax, nkd, h["ix_start"])
        barrier()
        t_sh = torch.tensor([(time.perf_counter() - t0) / nrep], dtype=torch.float64, device="cuda")
        dist.all_reduce(t_sh, op=dist.ReduceOp.MAX)
        tm_sh = ctx.timing()
        if rank == 0:
            ref = dc.spectra(h["k"], o, ells, kmin, kmax, nkd, h["ix_start"])
            t0 = time.perf_counter(); ref = dc.spectra(h["k"], o, ells, kmin, kmax, nkd, h["ix_start"]); t_one = time.perf_counter() - t0
            dev = max(float(np.abs(outs[i] / ref[i] - 1).max()) for i in (0, 2))
            strong = {"workload": "ONE cosmology of the C3-value workload, k-modes and multipoles sharded over the GPUs (bolt_spectra_sharded: "
                                  "K1 on a cyclic shard of the descending-k order, ncclAllGather of S_T/S_P, K2 on every N-th multipole, one "
                                  "ncclAllReduce of C_l)", "n_gpus": world, "ms_per_spectrum_set": 1e3 * float(t_sh.item()),
                      "kmode_solves_per_s": NK / float(t_sh.item()), "single_gpu_ms_same_run": 1e3 * t_one,
                      "speedup_vs_one_gpu": t_one / float(t_sh.item()), "efficiency": t_one / float(t_sh.item()) / world,
                      "kernel_ms_rank0": {"hierarchy": tm_sh["hierarchy_ms"], "projection": tm_sh["project_ms"], "total": tm_sh["total_ms"]},
                      "max_rel_diff_tt_ee_vs_single_gpu": dev,
                      "limiter": "the longest k-mode of a shard (serial ODE steps) and the l-independent part of K2 (dense-k source "
                                 "interpolation, done by every rank), not the collectives (64 MB all-gather, 60 KB all-reduce)"}
        # P(k) of ONE cosmology sharded the same way (bolt_plin_sharded: one all-gather, no reduction)
        kp = np.geomspace(10 * h["bg"].H0, 5000 * h["bg"].H0, 500)
        if strong is not None or rank != 0:
            op = abi.make_opts(50, 50, 20, reltol=1e-5, abstol=ABSTOL)
            dc.plin_sharded(kp, op)
            barrier(); t0 = time.perf_counter(); pks = dc.plin_sharded(kp, op); barrier()
            t_pk = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            dist.all_reduce(t_pk, op=dist.ReduceOp.MAX)
            if rank == 0:
                t0 = time.perf_counter(); pk1 = dc.plin(kp, op); t_pk1 = time.perf_counter() - t0
                strong["plin_sharded"] = {"workload": "plin, 500 log-spaced modes, n = 473, sharded over the GPUs (K1 + epilogue per shard, one ncclAllGather)",
                                          "ms": 1e3 * float(t_pk.item()), "single_gpu_ms_same_run": 1e3 * t_pk1,
                                          "bit_identical_to_single_gpu": bool(np.array_equal(pks[0], pk1[0])),
                                          "note": "500 modes are less than one wave of a single GPU (1184 warp slots): the time is the longest mode's serial "
                                                  "ODE steps on either side, so sharding buys nothing until n_k exceeds a wave"}
        barrier()
        ctx.comm_free()

    times = torch.tensor([wall, wall_e2e, dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    wall_max, e2e_max, dev_max = [float(v) for v in times.tolist()]
    if rank == 0:
        total_solves = NK * args.steps * world
        value = total_solves / wall_max
        fl = nstep_tot * f_step(n)        # accepted + rejected step attempts (bolt_spectra returns both counts)
        roof = {"kernel": "hierarchy_kernel_t<Trunc<8,8,10,15,19>, 4> (K1, dominant: 75 % of the step)", "bound": "fp64", "achieved": fl / (k1_ms * 1e-3) / 1e12, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": fl / (k1_ms * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                "traffic": 4.5e5, "traffic_note": "NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of ONE ncu --set full capture of this kernel "
                                                  "on the same 2000-mode grid (profiles/r2/k1_warp_lockstep_2000modes.md, round 2): 295 KB + 156 KB per launch -- "
                                                  "K1 never touches HBM in steady state",
                "peak_source": "DFMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 figure)",
                "note": "algorithmic flop = step attempts (accepted + rejected) x (182 n + 3300), n = 197 (DESIGN.md); stall analysis in profiles/"}
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm_peak = 6650.0
        nell = len(ells)
        k2_bytes = (2 * 799 * 4999 * 8) + nell * 5003 * 8      # one read of the dense source grids + the spline tables
        k2_terms = 2.0 * nell * 4999 * 799
        roof_k2 = {"kernel": "project_kernel (K2)", "bound": "hbm", "achieved": k2_bytes * args.steps / (k2_ms * 1e-3) / 1e9, "peak": hbm_peak,
                   "unit": "GB/s", "frac": k2_bytes * args.steps / (k2_ms * 1e-3) / 1e9 / hbm_peak, "traffic": 2.95e8,
                   "traffic_note": "NOT measured in this run: one ncu --set full capture of the shipped kernel (profiles/r2/k2_projection_final.md): 291 MB read + 4.4 MB written per launch",
                   "fp64_tflops": 25.0 * k2_terms / 2 * args.steps / (k2_ms * 1e-3) / 1e12,
                   "note": "K2 is FP64/shared-memory-gather bound, not HBM bound (SURVEY 8d): both figures reported"}
        line = {"metric": "kmode_hierarchy_solves_per_s", "value": value, "unit": "k-mode solves/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * wall_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "spectra_per_s": args.steps * world / wall_max, "device_ms_per_step": dev_max / args.steps,
                "kernel_ms_per_step": {"hierarchy": k1_ms / args.steps, "bessel_tables": bes_ms / args.steps, "projection": k2_ms / args.steps},
                "kernel_ms_note": "bessel_tables is the elapsed time of the second stream, enqueued behind K1's launch to fill its tail: it is "
                                  "mostly waiting for SMs (the two table kernels take 3.7 + 3.0 ms alone) and overlaps hierarchy",
                "ode_step_attempts_per_solve": nstep_tot / (NK * args.steps), "failed_modes": bad,
                "roofline": roof, "roofline_k2": roof_k2, "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": total_solves / e2e_max, "unit": "k-mode solves/s",
                        "h2d_bytes_per_step": int(hcos[0]["hc"].tables.nbytes + hcos[0]["hc"].scalars.nbytes + 2 * 15 * 8 + NK * 8 + nell * 4 + 2 * NK * 4),
                        "d2h_bytes_per_step": int(3 * nell * 8 + NK * 4 + NK * 8)}}
        if not args.no_cpu_baseline and world == 1:
            cb, _ = cpu_sample(hcos[0])
            line["cpu_baseline"] = cb
        # the extra arms must never cost the headline line
        def guarded(fn, *a):
            try:
                return fn(*a)
            except Exception as e:          # noqa: BLE001 -- reported in the JSON line instead of aborting the benchmark
                return {"error": f"{type(e).__name__}: {e}"[:300]}
        if not args.no_gradients and world == 1:
            line["gradients"] = guarded(gradient_arm, ctx, ells, not args.no_cpu_baseline)
        if strong is not None:
            line["strong_scaling"] = strong
        if world == 1:
            line["plin"] = guarded(plin_arm, ctx, hcos[0], dcs[0])
            line["batch"] = guarded(batch_arm, ctx, hcos, dcs, ells)
            line["hostgen"] = guarded(hostgen_arm, local_rank)
            line["params_to_spectra"] = guarded(params_to_spectra_arm, ctx, local_rank, ells)
            line["filon"] = guarded(filon_arm, local_rank)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

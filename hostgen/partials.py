"""Harness: host tables WITH parameter partials for the gradient tests and benchmark (the Julia host gets them for free from
ForwardDiff.Dual parameters; this Python generator has no AD and differentiates the whole host pipeline by central differences)."""
from bolt_b200 import abi
from .background import Background
from .recfast import RECFAST, IonizationHistory


def host_cosmo_with_partials(par, names, rel_step=1e-3, x_grid=None):
    """Pack Background + IonizationHistory WITH partials d/d(par.<name>) for the device (nd = 1 + len(names)).

    In Julia the partials come for free: Background / RECFAST run on ForwardDiff.Dual parameters and the spline
    coefficient arrays are Vector{Dual}.  This Python host mirror has no AD, so it differentiates the host tables
    by central differences of the whole host pipeline (2 extra host runs per parameter); what the device then does
    with those partials (K1, K2, plin) is exact forward-mode propagation."""
    def host(p):
        bg = Background(p) if x_grid is None else Background(p, x_grid=x_grid)
        ih = IonizationHistory(RECFAST(bg, OmegaB=p.Ω_b, Yp=p.Y_p, OmegaG=p.Ω_r), p, bg)
        return abi.HostCosmo.from_host(p, bg, ih), bg, ih
    base, bg, ih = host(par)
    pm, steps = [], []
    for nm in names:
        v = getattr(par, nm)
        dlt = rel_step * (abs(v) if v != 0 else 1.0)
        if nm in ("A", "n"):      # enter only through the primordial weight (spectra.jl:92): no host re-run needed
            hi = abi.HostCosmo(base.scalars.copy(), base.quad_pts, base.quad_wts, base.tables, base.x0, base.dx)
            lo = abi.HostCosmo(base.scalars.copy(), base.quad_pts, base.quad_wts, base.tables, base.x0, base.dx)
            hi.scalars[abi.S[nm], 0] = v + dlt; lo.scalars[abi.S[nm], 0] = v - dlt
            pm.append((hi, lo)); steps.append(dlt)
        else:
            pm.append((host(par.replace(**{nm: v + dlt}))[0], host(par.replace(**{nm: v - dlt}))[0])); steps.append(dlt)
    return abi.HostCosmo.with_partials(base, pm, steps), base, bg, ih, pm, steps

"""Static check of the Julia shim (Julia is not installed here, so the shim cannot run): every `ccall` in
bolt.jl_b200/julia/BoltCUDA.jl is parsed and its return type and argument tuple are compared, position by position, with the
prototype of the same function in include/bolt_cuda.h; the two mirrored structs are compared field by field."""
import os
import re

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "bolt_cuda.h")
SHIM = os.path.join(ROOT, "bolt.jl_b200", "julia", "BoltCUDA.jl")

# C parameter type -> the Julia ccall types that may stand for it
C2J = {
    "int": {"Cint"}, "int32_t": {"Int32", "Cint"}, "double": {"Cdouble", "Float64"},
    "bolt_ctx*": {"Ptr{Cvoid}"}, "const bolt_ctx*": {"Ptr{Cvoid}"}, "bolt_ctx**": {"Ref{Ptr{Cvoid}}"},
    "bolt_cosmo*": {"Ptr{Cvoid}"}, "const bolt_cosmo*": {"Ptr{Cvoid}"}, "bolt_cosmo**": {"Ref{Ptr{Cvoid}}"},
    "const bolt_cosmo* const*": {"Ptr{Ptr{Cvoid}}"},
    "const bolt_cosmo_desc*": {"Ref{CosmoDesc}"}, "const bolt_opts*": {"Ref{Opts}"},
    "const double*": {"Ptr{Float64}"}, "double*": {"Ptr{Float64}"},
    "const int32_t*": {"Ptr{Int32}"}, "int32_t*": {"Ptr{Int32}"}, "int64_t*": {"Ptr{Int64}"},
    "void*": {"Ptr{Cvoid}"}, "const void*": {"Ptr{Cvoid}"},
    "bolt_moment_table*": {"Ptr{Cvoid}"}, "bolt_moment_table**": {"Ref{Ptr{Cvoid}}"}, "float*": {"Ptr{Cfloat}"},
}
RET = {"int": "Cint", "const char*": "Cstring", "void": "Cvoid"}


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({":
            depth += 1
        elif ch in ")}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def c_prototypes():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int|void|const char\*)\s+(bolt_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        params = []
        for a in split_top(" ".join(args.split())):
            if a == "void":
                continue
            m = re.match(r"(.*?)(\b[A-Za-z_][A-Za-z_0-9]*)?$", a)      # strip the parameter name
            ty = re.sub(r"\s*\*\s*", "* ", m.group(1)).strip() if m.group(2) and m.group(1).strip() else a
            ty = re.sub(r"\s+", " ", ty).replace("* *", "**").replace(" *", "*").strip()
            ty = ty.replace("* const*", "* const*")
            params.append(ty)
        protos[name] = (ret, params)
    return protos


def julia_ccalls():
    src = open(SHIM).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(bolt_[a-z_0-9]+), lib\),\s*([A-Za-z]+),\s*\(", src):
        i = m.end(); depth = 1; j = i
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0); j += 1
        tup = src[i:j - 1]
        calls.append((m.group(1), m.group(2), [t for t in split_top(" ".join(tup.split())) if t]))
    return calls


def test_every_ccall_matches_the_header():
    protos = c_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 12
    used = set()
    for name, ret, jargs in calls:
        assert name in protos, f"{name} is not declared in bolt_cuda.h"
        cret, cargs = protos[name]
        used.add(name)
        assert RET[cret] == ret, (name, ret, cret)
        assert len(jargs) == len(cargs), (name, len(jargs), len(cargs), jargs, cargs)
        for pos, (j, c) in enumerate(zip(jargs, cargs)):
            assert c in C2J, (name, pos, c)
            assert j in C2J[c], f"{name} argument {pos}: Julia {j} vs C {c}"
    # the shim binds the whole hot path and the multi-GPU entry points
    assert {"bolt_init", "bolt_cosmo_upload", "bolt_cosmo_free", "bolt_solve", "bolt_project", "bolt_plin", "bolt_spectra_batch",
            "bolt_spectra_sharded", "bolt_comm_init", "bolt_comm_unique_id", "bolt_state_dim", "bolt_last_error"} <= used


def c_struct_fields(name):
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        ty, names = re.match(r"((?:const )?[a-z0-9_]+\*?)\s*(.*)", decl).groups()
        for nm in names.split(","):
            fields.append((ty.strip(), nm.strip()))
    return fields


def julia_struct_fields(name):
    src = open(SHIM).read()
    body = re.search(r"struct %s\n(.*?)\nend" % name, src, flags=re.S).group(1)
    return [(t, n) for n, t in re.findall(r"([A-Za-z_0-9]+)::([A-Za-z0-9{}]+)", body)]


def test_struct_mirrors_match_field_by_field():
    jt = {"int32_t": "Int32", "int64_t": "Int64", "double": "Float64", "const double*": "Ptr{Float64}"}
    for cname, jname in (("bolt_cosmo_desc", "CosmoDesc"), ("bolt_opts", "Opts")):
        cf, jf = c_struct_fields(cname), julia_struct_fields(jname)
        assert [n for _, n in cf] == [n for _, n in jf], (cname, cf, jf)
        assert [jt[t] for t, _ in cf] == [t for t, _ in jf], (cname, cf, jf)
    src = open(SHIM).read()
    hdr = open(HEADER).read()
    assert int(re.search(r"const ABI_VERSION = (\d+)", src).group(1)) == int(re.search(r"#define BOLT_ABI_VERSION (\d+)", hdr).group(1))


def test_shim_adds_methods_for_the_reference_signatures():
    """The reference's exported entry points of this path (src/Bolt.jl:8-9) each have a device method."""
    src = open(SHIM).read()
    for pat in (r"function boltsolve\(h::Hierarchy\{T,Device\}", r"function boltsolve_rsa\(h::Hierarchy\{T,Device\}",
                r"source_grid\(𝕡::AbstractCosmoParams, bg, ih, k_grid, dev::Device", r"source_grid_P\(𝕡::AbstractCosmoParams, bg, ih, k_grid, dev::Device",
                r"cltt\(ℓ::Int, 𝕡::AbstractCosmoParams, bg, ih, sf::DeviceSourceGrid\)", r"clte\(ℓ::Int,", r"clee\(ℓ::Int,",
                r"cltt\(ℓ⃗::AbstractVector,", r"cltt\(ℓ, s::DeviceSourceGrid, kgrid,", r"clte\(ℓ, s::DeviceSourceGrid, sP::DeviceSourceGrid, kgrid,",
                r"clee\(ℓ, sP::DeviceSourceGrid, kgrid,", r"function plin\(ks::AbstractVector,", r"plin\(k::Real, 𝕡::AbstractCosmoParams, bg, ih, dev::Device",
                r"function Bolt.sph_bessel_interpolator\(dev::Device, ν::Int, order, kη_min, kη_max, N::Int; weniger_cut=50\)",
                r"function Bolt.integrate_sph_bessel_filon\(f::AbstractVector,", r"Bolt.integrate_sph_bessel_filon\(f::Real,",
                r"Bolt.getnu\(t::DeviceMomentTable\)", r"Bolt.getorder\(t::DeviceMomentTable\)", r"function hostgen_batch\(pars::AbstractVector"):
        assert re.search(pat, src), pat

"""The C-ABI library loads and exports every symbol include/bolt_cuda.h declares; struct layouts match ctypes."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "bolt_cuda.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bolt_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from bolt_b200 import capi
    lib = C.CDLL(capi.LIB_PATH)
    names = declared_functions()
    assert {"bolt_init", "bolt_solve", "bolt_project", "bolt_spectra", "bolt_plin", "bolt_cosmo_upload"} <= set(names)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.bolt_abi_version() == 3
    assert lib.bolt_state_dim(8, 8, 10, 15) == 197 and lib.bolt_state_dim(50, 50, 20, 15) == 473


def test_struct_layout_matches_header(tmp_path):
    from bolt_b200 import abi
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "bolt_cuda.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %d %d\\n", sizeof(bolt_cosmo_desc), offsetof(bolt_cosmo_desc,x0),'
                    'offsetof(bolt_cosmo_desc,scalars), offsetof(bolt_cosmo_desc,tables), sizeof(bolt_opts), offsetof(bolt_opts,reltol),'
                    'offsetof(bolt_opts,max_steps), offsetof(bolt_opts,ix_first), (int)BOLT_NSCALARS, (int)BOLT_NTABLES);return 0;}\n')
    exe = tmp_path / "layout"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    D, O = abi.CosmoDesc, abi.Opts
    want = [C.sizeof(D), D.x0.offset, D.scalars.offset, D.tables.offset, C.sizeof(O), O.reltol.offset, O.max_steps.offset,
            O.ix_first.offset, abi.NSCALARS, abi.NTABLES]
    assert got == want


def test_no_cpu_fallback_without_device():
    """The product path must fail loudly when there is no usable GPU."""
    import torch
    from bolt_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.BoltError):
        capi.Context(0)
    import bolt_b200 as B
    with pytest.raises(capi.BoltError):
        B.default_context()
    # the context-free entry points (batched input tables, Bessel moments / Filon rule) select a device themselves: same rule
    import numpy as np
    import bolt_b200.bessel as BM
    with pytest.raises(capi.BoltError):
        capi.hostgen_batch([B.CosmoParams()])
    with pytest.raises(capi.BoltError):
        BM.sph_bessel_interpolator(3, 3, 0.0, 100.0, 1000)
    with pytest.raises(capi.BoltError):
        BM.sph_j_moment_asymp(200.0, 2, 0)
    with pytest.raises(capi.BoltError):
        BM._moments(2, [0, 1, 2], BM.SMALL, np.array([1.0, 10.0]))


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "bolt.jl_b200")
    pat = re.compile(r"(^|\s)(import\s+oracle|from\s+oracle)|oracle/|libbolt_oracle|OracleCosmo")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert not pat.search(txt), os.path.join(d, f)
                # run-time binding is allowed for NCCL only (bolt_comm_*): every library name handed to dlopen is an NCCL name
                if "dlopen" in txt:
                    names = re.search(r"const char\* names\[\] = \{([^}]*)\}", txt)
                    assert names and all("nccl" in lit.lower() for lit in re.findall(r'"([^"]*)"', names.group(1))), os.path.join(d, f)
                    assert len(re.findall(r"\bdlopen\(", txt)) == 1, os.path.join(d, f)
    # and bench.py touches it only inside its CPU-baseline legs: functions named cpu_* (the reported cpu_baseline and the
    # --impl reference arm both go through them), never at module level or inside the GPU arms
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    found = []

    def visit(node, fn):
        for ch in ast.iter_child_nodes(node):
            name = ch.name if isinstance(ch, (ast.FunctionDef, ast.AsyncFunctionDef)) else fn
            if isinstance(ch, ast.ImportFrom) and (ch.module or "").split(".")[0] == "oracle":
                found.append(fn)
            if isinstance(ch, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in ch.names):
                found.append(fn)
            visit(ch, name)
    visit(tree, None)
    assert found and all(fn is not None and fn.startswith("cpu_") for fn in found), found


def test_shard_plan_partitions_the_modes_in_work_order():
    """bolt_shard_plan is host-only: cyclic shards of the descending-k order, disjoint and complete, equal to the Python mirror."""
    import numpy as np
    from bolt_b200 import capi
    from bolt_b200.parallel import k_shard
    rng = np.random.default_rng(5)
    k = rng.random(37) + 0.1
    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            idx = capi.shard_plan(k, r, world)
            assert np.all(np.diff(k[idx]) < 0)                      # work order: longest solves (largest k) first
            assert np.array_equal(np.sort(idx), k_shard(k, r, world))
            seen.append(idx)
        assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(len(k)))

"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/mdb200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, load_pkg


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mdb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mdb_[A-Za-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    m = load_pkg()
    L = m.load_library()
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    # and the python mirror binds exactly the declared surface
    assert sorted(m.EXPORTS) == syms


def test_abi_version_and_default_params():
    m = load_pkg()
    L = m.load_library()
    assert L.mdb_abi_version() == 1
    p = m.default_params()
    # initParameter defaults, reference common/parameter.c:16-51
    assert (p.nx, p.ny, p.nz) == (32, 32, 32)
    assert p.ntimes == 200 and p.nstat == 100 and p.reneigh_every == 20 and p.half_neigh == 0
    assert p.dt == 0.005 and p.cutforce == 2.5 and p.skin == 0.3 and p.temp == 1.44
    assert p.rho == 0.8442 and p.epsilon == 1.0 and p.sigma == 1.0 and p.mass == 1.0
    assert (p.pbc_x, p.pbc_y, p.pbc_z) == (1, 1, 1) and p.ntypes == 1
    assert p.precision == m.DP and p.layout == m.AOS and p.force_field == m.FF_LJ


def test_params_struct_size_matches_header():
    # the ctypes mirror must have the C struct's size: compile a one-liner against the header
    import subprocess
    import tempfile
    m = load_pkg()
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include "mdb200.h"\nint main(void){printf("%zu\\n", sizeof(mdb_params));return 0;}\n')
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        n = int(subprocess.check_output([exe]).decode())
    assert n == C.sizeof(m.Params)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not fall back to anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = load_pkg()
    with pytest.raises(m.MdbError) as e:
        m.Simulation()
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's cpu_baseline leg may import, link or execute oracle/."""
    pk = os.path.join(ROOT, "md-bench_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|oracle/|portbind|refbind|libmdoracle|libmdref|vl_oracle|ovl_")
    for dp, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f)).read()
                m = bad.search(txt)
                assert m is None, (os.path.join(dp, f), m.group(0))


def test_no_cpu_fallback_clusterpair_and_decomposition():
    """the other two handles of the C ABI fail just as loudly without a CUDA device"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = load_pkg()
    with pytest.raises(m.MdbError) as e:
        m.ClusterSimulation()
    assert "CUDA" in str(e.value)
    with pytest.raises(m.MdbError) as e:
        m.Decomposition(m.default_params(nx=8, ny=8, nz=8), (2, 1, 1))
    assert "CUDA" in str(e.value)


def test_every_runtime_option_is_documented_in_the_header():
    """mdb_setOption / mdb_cp_setOption / mdb_dd_setOption names accepted by the library (the strcmp chains in csrc/) must
    all appear, quoted, in include/mdb200.h -- the header is the only place a driver maintainer learns about them."""
    hdr = open(os.path.join(ROOT, "include", "mdb200.h")).read()
    csrc = os.path.join(ROOT, "md-bench_b200", "csrc")
    names = set()
    for f in ("sim_impl.cu", "cp_sim.cu", "dd_group.cuh"):
        names |= set(re.findall(r'!strcmp\(name, "([a-z_0-9]+)"\)', open(os.path.join(csrc, f)).read()))
    assert {"fuse_force", "xy_gather", "lazy_ops", "halo_push"} <= names
    missing = sorted(n for n in names if '"%s"' % n not in hdr)
    assert not missing, "options not documented in include/mdb200.h: %s" % missing


def test_traffic_json_names_the_kernels_that_run():
    """profiles/traffic.json feeds roofline.traffic; bench.py only accepts an entry captured from the fused kernels that
    mdb_run / mdb_cp_run actually launch (names as ncu prints them)."""
    import json
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    vl, cp = t["verletlist/dp/128"], t["clusterpair/sp/128"]
    assert "k_force_lj_full_fi" in vl["kernel"] and vl["bytes"] > 3.0e9
    assert "k_cp_force_lj_sp_duo<4, 1, 0>" in cp["kernel"] and cp["bytes"] > 0.9e9
    assert vl.get("commit") and cp.get("commit"), "every capture is stamped with the commit it was taken at"
    for e in (vl, cp):
        assert os.path.exists(os.path.join(ROOT, e["source"].split(" ")[0]))

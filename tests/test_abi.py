"""The C-ABI boundary without a GPU: the library builds, loads, exports exactly the symbols
include/flexigal_gpu.h declares, the Python host layer is the only caller, and nothing in the product
reaches for the oracle or a CPU fallback."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "flexigal_gpu.h")


def declared_symbols():
    txt = open(HEADER).read()
    return sorted(set(re.findall(r"FG_API\s+[\w\s\*]+?\b(fg_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(fg):
    from flexigal_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the extension first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SYMBOLS) == decl, "flexigal_b200/_lib.py must bind exactly the header's symbols"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (fg_\w+)", out)))
    assert exported == decl, "the .so exports symbols the header does not declare (or vice versa)"
    abi, sm = ctypes.c_int(), ctypes.c_int()
    assert lib.fg_version(ctypes.byref(abi), ctypes.byref(sm)) == 0 and sm.value == 100


def test_built_for_sm_100a_only():
    from flexigal_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_and_product_does_not_touch_the_oracle(fg):
    import torch
    pkg = os.path.join(ROOT, "flexigal_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "efg_oracle" not in txt and "import oracle" not in txt, f"{f} references the oracle"
    if not torch.cuda.is_available():
        with pytest.raises(fg.FlexiGalGPUError, match="no CUDA device|CPU path"):
            fg.Context(0)
        model = fg.create_model((1.0, 1.0), (3, 3))
        with pytest.raises(fg.FlexiGalGPUError):
            fg.SHAPE_FUN(np.array([[0.5, 0.5, 1.0, 1.0]]), model.x, fg.Influence_Domains(model, 1.5))


def test_operator_struct_layout_matches_header():
    from flexigal_b200 import _lib
    assert ctypes.sizeof(_lib.fg_operator) == 4 + 4 + 32 + 8 + 8
    assert _lib.fg_operator.coef.offset == 40 and _lib.fg_operator.coef_stride.offset == 48


def test_host_geometry_matches_oracle_setup_bitwise(fg, oracle):
    """flexigal_b200/geometry.py (what crosses the ABI in tests/bench) == the oracle's restatement of
    Meshing.jl / Geometry.jl / Integration.jl / Influence_Domains, bit for bit."""
    for dom, div, deg in [((2.0, 1.0), (7, 5), 2), ((1.0, 2.0, 1.5), (5, 4, 3), 3), ((1.0, 1.0), (4, 4), 4)]:
        om, pm = oracle.create_model(dom, div), fg.create_model(dom, div)
        assert np.array_equal(om.x, pm.x) and np.array_equal(om.conn, pm.conn)
        tags = ["Domain", "Left", "Right", "Top", "Bottom", "Boundary"] + (["Back", "Front"] if len(dom) == 3 else [])
        for tag in tags:
            a, _ = oracle.background_integration(om, tag, deg)
            b = fg.BackgroundIntegration(fg.IntegrationSet(fg.Triangulation(pm, tag), deg)).gs
            assert np.array_equal(a, b), tag
        for dmax, shape in [(1.5, "rectangular"), ([1.2, 1.7, 1.4][:len(dom)], "rectangular"), (2.0, "circular")]:
            assert np.array_equal(oracle.influence_domains(om, dmax, shape), fg.Influence_Domains(pm, dmax, shape))
    with pytest.raises(ValueError, match="Unknown shape of influence domain"):
        fg.Influence_Domains(pm, 1.5, "triangular")
    with pytest.raises(ValueError, match="ngpts"):
        fg.pgauss(5)


def test_operator_descriptors_marshal_like_the_header(fg):
    op = fg.Elasticity(2.0, 3.0, D=3).c_struct(3, 10)
    assert (op.kind, op.D, op.c[0], op.c[1]) == (4, 3, 2.0, 3.0) and not op.coef
    C = np.arange(16.0).reshape(4, 4)
    g = fg.Generic(C)
    op = g.c_struct(3, 10)
    assert op.kind == 0 and op.coef_stride == 0 and op.coef
    op = fg.Generic(np.zeros((10, 4, 4))).c_struct(3, 10)
    assert op.coef_stride == 16
    src = fg.Source(np.arange(10.0))
    op = src.c_struct(3, 10)
    assert op.kind == 0 and op.coef_stride == 4 and np.array_equal(src._keep.reshape(10, 4)[:, 0], np.arange(10.0))
    assert fg.Source(5.0).c_struct(2, 3).kind == 1
    probed = fg.Closure(lambda sa, sb: 2.0 * float(sa.dphi @ sb.dphi) + sa.phi * sb.phi)
    coef, stride = probed.coefficients(2, 5)
    assert stride == 0 and np.array_equal(coef.reshape(3, 3), np.diag([1.0, 2.0, 2.0]))


def test_committed_launch_list_parses_and_names_the_step_kernels():
    """profiles/r02_launches_n100_final.csv is the evidence bench.py's roofline.traffic (profiles/traffic.json) comes from: the
    summary tool must read it and agree with the JSON; round 1's list keeps parsing too."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import launch_summary
    import json
    traffic = json.load(open(os.path.join(root, "profiles", "traffic.json")))
    table = launch_summary.summarise(os.path.join(root, "profiles", "r02_launches_n100_final.csv"))
    rows = {l.split("`")[1]: [c.strip() for c in l.split("|")[2:-1]] for l in table.splitlines()[2:]}
    for name in ("k_imls_c", "k_bilinear_pts", "k_nb_cluster"):
        hit = [v for k, v in rows.items() if name in k]
        assert hit, name
        rd, wr = float(hit[0][2]), float(hit[0][3])
        assert abs((rd + wr) * 1e9 - traffic[name]) <= 0.02 * traffic[name], (name, rd + wr, traffic[name])   # bench.py's roofline.traffic
    table = launch_summary.summarise(os.path.join(root, "profiles", "r01_launches_n100_final.csv"))
    rows = {l.split("`")[1]: [c.strip() for c in l.split("|")[2:-1]] for l in table.splitlines()[2:]}
    for name in ("k_imls", "k_bilinear_reg", "k_nb_cluster"):
        hit = [v for k, v in rows.items() if name in k]
        assert hit, name
        launches, avg_ms, rd, wr = hit[0][0], float(hit[0][1]), float(hit[0][2]), float(hit[0][3])
        assert int(launches) >= 2 and avg_ms > 0.5 and rd + wr > 1.0


def test_integration_shim_calls_only_declared_symbols_and_is_replayed():
    """Every `ccall` of INTEGRATION.md's Julia shim names a symbol the header declares, and the ctypes replay test
    (tests/test_gpu_round2.py) goes through every one of them."""
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    called = sorted(set(re.findall(r"ccall\(\(:(fg_\w+), LIB\)", md)))
    assert len(called) >= 14
    decl = set(declared_symbols())
    assert set(called) <= decl, set(called) - decl
    replay = open(os.path.join(ROOT, "tests", "test_gpu_round2.py")).read()
    body = replay[replay.index("def test_linear_problem_call_sequence_replayed_through_ctypes"):]
    body = body[:body.index("\n@pytest.mark.parametrize")]
    missing = [name for name in called if f"lib.{name}(" not in body]
    assert not missing, f"INTEGRATION.md calls {missing}, which the replay test never makes"


def test_every_kernel_selection_switch_is_registered_and_documented():
    """fg_set_option refuses unknown keys, so every switch a kernel translation unit reads with fg_opt() has to be in its list of known keys
    (release builds; `debug_skip` exists in tuning builds only) -- and the switches INTEGRATION.md tells a shim about have to exist."""
    src = os.path.join(ROOT, "flexigal_b200", "csrc")
    used = set()
    for f in os.listdir(src):
        if f.endswith((".cu", ".cuh")):
            used |= set(re.findall(r'fg_opt\(\s*\w+\s*,\s*"(\w+)"', open(os.path.join(src, f)).read()))
    api = open(os.path.join(src, "fg_api.cu")).read()
    block = api[api.index("static const char *const known[]"):]
    block = block[:block.index("nullptr};")]
    known = set(re.findall(r'"(\w+)"', block))
    assert used - known - {"debug_skip"} == set(), f"read by a kernel file but refused by fg_set_option: {sorted(used - known)}"
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for key in re.findall(r'fg_set_option\(ctx, "(\w+)"', doc):
        assert key in known, key
    for key in ("imls_grid", "pattern_tables", "cluster_lidx_eager", "lazy_shapes", "xchg_direct"):
        assert key in known and key in doc, key

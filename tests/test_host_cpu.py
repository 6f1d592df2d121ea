"""Host-side logic of the product package, checked on the CPU: factories and node orders against
the golden tree dump, the tree compiler through a NumPy emulation of the kernel's walk, the
launch-geometry (broadcast mask) logic, the C ABI surface, and the "no CPU fallback" errors."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import ultrasphere_b200 as us
from oracle import vks_oracle as vo
from tests.helpers import emulate_from_cartesian, emulate_to_cartesian, load_trees, make
from ultrasphere_b200 import _native
from ultrasphere_b200._plan import plan_for
from ultrasphere_b200._transform import _Geometry

TREES = load_trees()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("spec", sorted(TREES))
def test_factories_match_reference_dump(spec):
    c = make(us, spec)
    g = TREES[spec]
    assert list(c.s_nodes) == g["s_nodes"]
    assert list(c.c_nodes) == g["c_nodes"]
    assert [list(e) for e in c.cos_edges] == g["cos_edges"]
    assert [list(e) for e in c.sin_edges] == g["sin_edges"]
    assert [[k, v.value] for k, v in c.branching_types.items()] == g["branching_types"]
    assert [[k, (None if v != v else int(v))] for k, v in c.S.items()] == g["S"]
    assert c.root == g["root"] and list(c.G.nodes) == g["G_nodes"]
    assert (c.s_ndim, c.c_ndim) == (g["s_ndim"], g["c_ndim"])
    if c.s_ndim > 0:
        assert repr(c) == g["repr"] and str(c) == g["repr"]
        assert c.branching_types_expression_str == g["expression_str"]
        assert c.root_index == 0


# reference tests/test_creation.py:11-21
@pytest.mark.parametrize("s_ndim", [0, 1, 2, 3, 6, 103])
def test_generate(s_ndim):
    assert us.create_random(s_ndim).s_ndim == s_ndim
    assert us.create_standard(s_ndim).s_ndim == s_ndim
    assert us.create_standard_prime(s_ndim).s_ndim == s_ndim


@pytest.mark.parametrize("q", [0, 1, 2, 5])
def test_generate_hopf(q):
    assert us.create_hopf(q).c_ndim == 2**q


def test_bad_expressions_raise_value_error():
    for expr in ("x", "ab", "aa", "b", "ca", "bq"):
        with pytest.raises(ValueError):
            us.create_from_branching_types(expr)
    with pytest.raises(ValueError):
        us.create_hopf(-1)


def test_identity_semantics_and_pickle():
    import pickle

    a, b = us.create_standard(3), us.create_standard(3)
    assert a != b and a == a and hash(a) == hash(a)  # identity of G, as in the reference
    c = pickle.loads(pickle.dumps(us.create_hopf(2)))
    assert repr(c) == "SphericalCoordinates(caa)" and c.c_nodes == [0, 1, 2, 3]
    assert us.get_child(a.G, "theta0", "sin") == "theta1" and us.get_parent(a.G, "theta1") == "theta0"
    assert us.get_parent(a.G, "theta0") is None
    assert a.is_c_keys_range and not a.is_s_keys_range
    assert abs(us.create_standard(2).surface_area() - 4 * np.pi) < 1e-12
    assert abs(us.create_standard(2).volume() - 4 * np.pi / 3) < 1e-12


PLAN_SPECS = [s for s in sorted(TREES) if TREES[s]["s_ndim"] > 0]


@pytest.mark.parametrize("spec", PLAN_SPECS)
def test_compiled_program_reproduces_oracle(spec):
    """The pre-order program, walked the way the kernels walk it, gives the oracle's numbers:
    bit-exact (same NumPy ufuncs, same association), so any mismatch is a tree-compiler bug."""
    c = make(us, spec)
    t = vo.make(spec)
    ct = plan_for(c)
    assert ct.n_nodes == c.s_ndim and ct.n_leaves == c.c_ndim
    rng = np.random.default_rng(5)
    for dt in (np.float64, np.float32):
        x = rng.uniform(-1, 1, (c.c_ndim, 33)).astype(dt)
        want = vo.from_cartesian(t, x)
        r, ang = emulate_from_cartesian(ct, list(x))
        assert np.array_equal(r, want["r"])
        for i, node in enumerate(c.s_nodes):
            assert np.array_equal(ang[i], want[node]), (spec, node)
        sph = {node: ang[i] for i, node in enumerate(c.s_nodes)} | {"r": r}
        cart = emulate_to_cartesian(ct, ang, r)
        wantc = vo.to_cartesian(t, sph)
        for i, leaf in enumerate(c.c_nodes):
            assert np.array_equal(cart[i], wantc[leaf]), (spec, leaf)
    # structure tables used for output shapes
    for li, leaf in enumerate(c.c_nodes):
        anc = []
        node = leaf
        while us.get_parent(c.G, node) is not None:
            node = us.get_parent(c.G, node)
            anc.append(c.s_nodes.index(node))
        assert list(ct.leaf_ancestors[li]) == anc


def test_plan_capacity_and_validation():
    with pytest.raises(ValueError):
        plan_for(us.create_standard(_native.US_MAX_NODES + 1))
    assert plan_for(us.create_standard(_native.US_MAX_NODES)).n_nodes == _native.US_MAX_NODES
    # a c-chain deeper than the deferred-branch stack is refused, not silently mis-run
    deep = "c" * (_native.US_MAX_STACK + 1) + "a" * (_native.US_MAX_STACK + 2)
    with pytest.raises(ValueError):
        plan_for(us.create_from_branching_types(deep))
    ok = "c" * _native.US_MAX_STACK + "a" * (_native.US_MAX_STACK + 1)
    assert plan_for(us.create_from_branching_types(ok)).max_stack == _native.US_MAX_STACK
    lib = _native.lib()
    plan = _native.UsPlan()
    kinds = (ctypes.c_uint8 * 2)(1, 1)  # "bb" never closes
    slots = (ctypes.c_int32 * 2)(0, 1)
    leaves = (ctypes.c_int32 * 3)(0, 1, 2)
    assert lib.us_plan_compile(2, kinds, slots, leaves, ctypes.byref(plan)) == -3
    assert b"close" in lib.us_last_error_string()


def test_library_exports_every_declared_symbol():
    """include/ultrasphere_b200.h is the contract: every US_API function must be in the .so."""
    with open(os.path.join(ROOT, "include", "ultrasphere_b200.h")) as fh:
        header = fh.read()
    declared = set(re.findall(r"US_API\s+[\w\s\*]+?\b(us_\w+)\s*\(", header))
    assert declared == set(_native.EXPORTED), declared ^ set(_native.EXPORTED)
    lib = _native.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.us_abi_version() == _native.US_ABI_VERSION
    assert b"sm_100a" in lib.us_version()
    assert ctypes.sizeof(_native.UsPlan) == 16 + 8 * _native.US_MAX_NODES
    assert lib.us_weighted_reduce_scratch_bytes(3) >= 3 * 8


def test_argument_errors_do_not_need_a_gpu():
    lib = _native.lib()
    assert lib.us_to_cartesian(None, 0, 0, 1, None, None, None, None, 0, None, None) == -1
    assert b"null plan" in lib.us_last_error_string()
    assert lib.us_weighted_reduce(7, 1, None, None, None, 1, None, 0, None, 0, None) == -1
    assert lib.us_separable_reduce(0, 0, None, None, None, None, None, None) == -2
    assert lib.us_quadrature_rules(_native.US_MAX_AXES + 1, None, None, None, None, None, None, None, None) == -2
    assert lib.us_quadrature_rules(1, None, None, None, None, None, None, None, None) == -1
    assert lib.us_quadrature_rules(0, None, None, None, None, None, None, None, None) == 0


def test_tuning_knobs_are_named_and_checked():
    """us_set_tuning is host-only state: every documented knob is accepted (0 = automatic restores the
    default), anything else is refused with US_EINVAL and a message naming it."""
    lib = _native.lib()
    for knob in (b"tile_groups", b"tile_bufs", b"tile_u", b"tile_max_rows", b"tiled_min_tiles", b"ring_u", b"ring_ctas", b"ring_slots"):
        assert lib.us_set_tuning(knob, 2) == 0, knob
    assert lib.us_set_tuning(b"tile_max_rows", 13) == 0 and lib.us_set_tuning(b"tiled_min_tiles", 8) == 0
    for knob in (b"tile_groups", b"tile_bufs", b"tile_u", b"ring_u", b"ring_ctas", b"ring_slots"):
        assert lib.us_set_tuning(knob, 0) == 0, knob
    assert lib.us_set_tuning(b"ring_w", 4) == -1
    assert b"ring_w" in lib.us_last_error_string()


def test_no_slot_is_released_before_its_loads_return():
    """Static check of the shipped SASS (tools/sass_release_scan.py): in every TMA-staged kernel a
    consumer's mbarrier arrive — the one that hands a shared-memory slot back to the producer — sits
    behind an instruction that reads the registers loaded from that slot, so in-order issue holds it
    until the loads have returned.  Without that, a refill was seen to overtake a queued load (the
    ring kernels' r row: a few dozen wrong points per 10^9, tools/stress.py)."""
    import shutil
    import subprocess
    import sys

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    done = subprocess.run([sys.executable, os.path.join(root, "tools", "sass_release_scan.py")], capture_output=True, text=True, timeout=600)
    assert done.returncode == 0, done.stdout[-3000:]
    assert "release-before-use: 0 of" in done.stdout and "from_cartesian_ring" in done.stdout and "wreduce_tma" in done.stdout
    # the checker itself: the pattern the ring kernels shipped with before the fix is flagged, the fixed one is not
    sys.path.insert(0, os.path.join(root, "tools"))
    try:
        from sass_release_scan import scan
    finally:
        sys.path.pop(0)
    head = "\tFunction : k\n"
    load = "        /*0010*/                   LDS.128 R16, [R10] ;\n        /*0020*/                   LDS.128 R20, [R10+0x1000] ;\n"
    arrive = "        /*0050*/              @!P1 SYNCS.ARRIVE.TRANS64.A1T0 RZ, [R11+URZ+0x100], RZ ;\n"
    use = "        /*0030*/                   LOP3.LUT R2, R16, R20, RZ, 0xfc, !PT ;\n"
    unrelated = "        /*0030*/                   IADD3 R2, PT, PT, R3, 0x1, RZ ;\n"
    assert scan(head + load + unrelated + arrive)[:2] == (2, 2)
    assert scan(head + load + use + arrive)[:2] == (0, 2)
    assert scan(head + load + "        /*0030*/                   MEMBAR.SC.CTA ;\n" + arrive)[:2] == (0, 2)


def test_rule_backend_switch_and_rule_parameters():
    from ultrasphere_b200._integral import _rule_parameters

    assert us.set_rule_backend("scipy") == "scipy"
    with pytest.raises(ValueError):
        us.set_rule_backend("numpy")
    with us.rule_backend("device"):
        assert us.set_rule_backend("device") == "device"
        with pytest.raises(RuntimeError, match="CUDA devices only"):
            us.roots(us.create_spherical(), 4, expand_dims_x=False, xp=torch, device="cpu")
        with pytest.raises(ValueError, match="positive integer"):
            us.roots(us.create_spherical(), 0, expand_dims_x=False, xp=torch, device="cuda")
    assert us.set_rule_backend("scipy") == "scipy"
    # (alpha, beta) are what the reference hands to roots_jacobi (_integral.py:88-103)
    c = us.create_from_branching_types("cbaa")  # theta0: c (cos->theta1: b, sin->theta2: a)
    got = {str(node): _rule_parameters(c, node) for node in c.s_nodes}
    S = {str(k): v for k, v in c.S.items()}
    assert got["theta0"][0] == "c" and got["theta0"][1:] == (S["theta1"] / 2, S["theta2"] / 2)
    kinds = sorted(v[0] for v in got.values())
    assert kinds == ["a", "a", "b", "c"] or kinds == ["a", "b", "c"]


def test_no_cpu_fallback():
    c = us.create_spherical()
    with pytest.raises(TypeError):
        c.from_cartesian(np.asarray([1.0, 2.0, 3.0]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c.from_cartesian(torch.ones(3, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c.to_cartesian({"theta": torch.ones(2), "phi": torch.ones(2)})
    with pytest.raises(KeyError):
        c.to_cartesian({"theta": torch.ones(2)})
    with pytest.raises(TypeError):
        us.roots(c, 4, expand_dims_x=True, xp=np)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        us.integrate(c, lambda s: s["theta"], False, 4, xp=torch, device="cpu")


def test_zero_angle_tree_passthrough():
    c = us.create_standard(0)
    x = torch.tensor([-2.0, 3.0])
    assert c.from_cartesian({"theta0": x})["r"] is x  # no abs, like the reference
    assert c.to_cartesian({"r": x})["theta0"] is x


def test_launch_geometry_masks():
    n = 5
    xs = [torch.zeros(n, 1, 1), torch.zeros(1, n, 1), torch.zeros(1, 1, 2 * n)]
    geo = _Geometry(list(xs) + [None], (n, n, 2 * n))
    assert geo.shape == [n, n, 2 * n] and geo.masks == [0b001, 0b010, 0b100, 0]
    # equal dense shapes collapse to one dimension
    same = [torch.zeros(2, 3, 4) for _ in range(3)]
    geo = _Geometry(same, (2, 3, 4))
    assert geo.shape == [24] and geo.masks == [1, 1, 1]
    # a 0-d operand next to dense ones is a pure broadcast
    geo = _Geometry([torch.zeros(2, 3), torch.zeros(())], (2, 3))
    assert geo.shape == [6] and geo.masks == [1, 0]
    # expanded (stride-0) operands are recognised as broadcasts without a copy
    base = torch.zeros(4, 1)
    geo = _Geometry([base.expand(4, 6), torch.zeros(4, 6)], (4, 6))
    assert geo.masks == [0b01, 0b11] and geo.tensors[0].data_ptr() == base.data_ptr()
    # a transposed view is densified (one copy), adjacent dims that all operands share merge
    t = torch.arange(12.0).reshape(3, 4).t()
    geo = _Geometry([t, torch.zeros(4, 3)], (4, 3))
    assert geo.shape == [12] and geo.tensors[0].is_contiguous()
    # more independent axes than the parameter block holds: densified instead of failing
    k = _native.US_MAX_DIMS + 1
    ops = [torch.zeros((1,) * i + (2,) + (1,) * (k - i - 1)) for i in range(k)]
    geo = _Geometry(ops, (2,) * k)
    assert geo.shape == [2**k] and all(m == 1 for m in geo.masks)


def test_custom_ops_trace_without_a_gpu():
    """torch.ops.ultrasphere_b200.* carry fake (shape-inferring) implementations, so FX /
    torch.compile tracing works on any host; the real implementations exist for CUDA only."""
    from torch._subclasses.fake_tensor import FakeTensorMode

    c = us.create_standard(3)
    h = us.plan_handle(c)
    assert us.plan_handle(c) == h and us.plan_handle(us.create_standard(3)) != h
    with FakeTensorMode():
        ang = [torch.empty(5, 1, device="cuda"), torch.empty(1, 7, device="cuda"), torch.empty(5, 7, device="cuda", dtype=torch.float64)]
        out = torch.ops.ultrasphere_b200.to_cartesian(ang, None, h)
        assert out.shape == (4, 5, 7) and out.dtype == torch.float64
        s = torch.ops.ultrasphere_b200.from_cartesian([torch.empty(9, device="cuda") for _ in range(4)], h)
        assert s.shape == (4, 9) and s.dtype == torch.float32
        v = torch.empty(4, 5, 3, device="cuda", dtype=torch.complex128)
        w = [torch.empty(4, device="cuda", dtype=torch.float64), torch.empty(5, device="cuda", dtype=torch.float64)]
        assert torch.ops.ultrasphere_b200.weighted_reduce(v, w).shape == (3,)
    with pytest.raises(NotImplementedError):
        torch.ops.ultrasphere_b200.from_cartesian([torch.zeros(3) for _ in range(4)], h)


def test_plain_integer_broadcast_shapes_matches_torch():
    import random

    from ultrasphere_b200._transform import broadcast_shapes

    rnd = random.Random(0)
    for _ in range(500):
        nd = rnd.randint(0, 4)
        base = [rnd.choice([0, 1, 2, 3, 5]) for _ in range(nd)]
        shapes = []
        for _ in range(rnd.randint(1, 5)):
            d = rnd.randint(0, nd)
            shapes.append(tuple(rnd.choice([1, base[nd - d + i]]) for i in range(d)))
        assert broadcast_shapes(*shapes) == tuple(torch.broadcast_shapes(*shapes)), shapes
    with pytest.raises(RuntimeError):
        broadcast_shapes((2, 3), (4, 3))


def test_header_is_plain_c_and_links(tmp_path):
    """include/ultrasphere_b200.h compiles as C99 and a C program can drive the library."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    _native.lib()  # make sure the library exists (conftest builds it when missing)
    exe = tmp_path / "abi_check"
    subprocess.run(
        ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
         os.path.join(root, "tests", "c_abi", "abi_check.c"), "-o", str(exe),
         "-L", os.path.dirname(_native.LIB_PATH), "-lultrasphere_b200", f"-Wl,-rpath,{os.path.dirname(_native.LIB_PATH)}"],
        check=True,
    )
    done = subprocess.run([str(exe)], capture_output=True, text=True)
    assert done.returncode == 0, (done.returncode, done.stdout, done.stderr)
    assert done.stdout.startswith("abi 1")


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the arm the driver runs next to ours) needs no GPU: exactly one
    JSON line on stdout with the contract's keys."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    done = subprocess.run(
        [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--points", "60000"],
        capture_output=True, text=True, cwd=root, timeout=300,
    )
    assert done.returncode == 0, done.stderr[-2000:]
    lines = [l for l in done.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches", "quadrature"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["gpu_launches"] == 0
    from oracle import refload

    # the unmodified reference where it can be loaded (the mount, or the copy oracle/make_ref.py ships), else the port
    assert line["cpu_baseline"]["kind"] == ("reference" if refload.available() else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["single_thread"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert abs(line["quadrature"]["reference_form"]["value"] - 29.449397825582118) < 1e-9


def test_roots_returns_fresh_tensors_every_call():
    """The host tables are cached (Golub–Welsch is slow); callers must get copies, never views of
    the cache: an in-place edit of a returned tensor must not change what the next call returns
    (the reference returns fresh arrays, _integral.py:108-109)."""
    import ultrasphere_b200 as us

    c = us.create_standard(3)
    xs, ws = us.roots(c, 6, expand_dims_x=False, xp=torch, device="cpu", dtype=torch.float64)
    keep_x = {k: v.clone() for k, v in xs.items()}
    keep_w = {k: v.clone() for k, v in ws.items()}
    for v in list(xs.values()) + list(ws.values()):
        v.mul_(0)
    xs2, ws2 = us.roots(c, 6, expand_dims_x=False, xp=torch, device="cpu", dtype=torch.float64)
    for k in c.s_nodes:
        assert torch.equal(xs2[k], keep_x[k]) and torch.equal(ws2[k], keep_w[k]), k
        assert float(ws2[k].abs().sum()) > 0


def test_reference_copy_recipe_ships_the_unmodified_modules(tmp_path, monkeypatch):
    """oracle/make_ref.py copies the four reference modules verbatim (sha256 in the manifest) and
    oracle/refload.py can load them from the copy when the mount is masked."""
    import hashlib
    import json
    import subprocess
    import sys

    from oracle import make_ref

    if not os.path.isfile(os.path.join(make_ref.SRC, make_ref.FILES[0])):
        pytest.skip("reference not mounted: the copy is made in the build container")
    assert make_ref.main()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    manifest = json.load(open(os.path.join(root, "oracle", "_ref", "MANIFEST.json")))
    for name, digest in manifest["files"].items():
        with open(os.path.join(make_ref.SRC, name), "rb") as fh:
            assert hashlib.sha256(fh.read()).hexdigest() == digest, name
    code = (
        "import numpy as np; from oracle import refload; ref = refload.load(); assert '_ref' in refload.source_dir(); "
        "c = ref.us.create_standard(3); s = c.from_cartesian(np.array([[0.25], [-0.5], [0.75], [-1.0]])); print(float(s['r'][0]))"
    )
    done = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, env=dict(os.environ, US_REF_IGNORE_MOUNT="1"))
    assert done.returncode == 0, done.stderr[-2000:]
    assert abs(float(done.stdout.strip()) - 1.3693063937629153) < 1e-15

"""Parity tests proper: the CUDA path (through the C ABI) against (a) the reference's recorded
outputs in tests/golden/ and (b) the CPU oracle on the same seeded inputs.

Bars
----
* index tables, permutation, nodes, packed matrices: bit-exact (np.array_equal)
* fnt (default and precomputation=False), dx, dxT, ifnt(exact_ifnt=True): bit-exact
* ifnt (default, FMA) : max|new-ref| / max|ref| <= 1e-13   (north star: 1e-12)
* eval (nested Horner): max|new-ref| / max|ref| <= 1e-12
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, case_kwargs, golden_cases, load_case

pytestmark = pytest.mark.gpu

TOL_IFNT_FMA = 1e-13
TOL_EVAL = 1e-12


def relmax(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def T():
    import lpfun_b200

    return lpfun_b200.Transform


def make(T, g, **kw):
    kwargs = case_kwargs(g)
    kwargs.update(report=False, precompilation=False)
    kwargs.update(kw)
    return T(**kwargs)


@pytest.mark.parametrize("name", golden_cases("newton"))
def test_golden_case(T, name):
    g = load_case(name)
    t = make(T, g, exact_ifnt=True)
    m = int(g["m"])
    # --- plan tables ---
    assert np.array_equal(t.multi_index_set, g["A"].astype(np.int64))
    assert np.array_equal(t.tube, g["T"])
    assert np.array_equal(t.nodes, g["x"]) and np.array_equal(t.leja_order, g["leja_order"])
    if bool(g["colex"]):
        assert np.array_equal(t.colex_order, g["P"])
    else:
        assert t.colex_order is None
    assert len(t) == int(g["N"])
    # --- fnt: bit exact, batched call == row by row ---
    got = t.fnt(g["values"])
    assert got.shape == g["fnt"].shape
    nbits = int(np.sum(got != g["fnt"]))
    assert nbits == 0, f"{nbits} coefficients differ, rel {relmax(got, g['fnt']):.2e}"
    assert np.array_equal(t.fnt(g["values"][0]), g["fnt"][0])
    assert np.array_equal(t.fnt(g["rnd"]), g["fnt_rnd"])
    # --- ifnt exact + FMA ---
    assert np.array_equal(t.ifnt(g["fnt"]), g["ifnt"])
    assert np.array_equal(t.ifnt(g["rnd"]), g["ifnt_rnd"])
    tf = make(T, g)  # default: FMA ifnt
    assert relmax(tf.ifnt(g["fnt"]), g["ifnt"]) <= TOL_IFNT_FMA
    # --- dx / dxT ---
    # p = inf, m > 1: the reference folds every row of dx from its last term down (molecules.py:290-294); so does the plan
    for key in g.files:
        if key.startswith("dx_"):
            _, i, k = key.split("_")
            out = t.dx(g["fnt"][0], int(i), int(k))
            assert np.array_equal(out, g[key]), key
        elif key.startswith("dxT_"):
            _, i, k = key.split("_")
            assert np.array_equal(t.dxT(g["rnd"], int(i), int(k)), g[key]), key
    # --- eval ---
    assert relmax(t.eval(g["fnt"][0], g["points"]), g["eval"]) <= TOL_EVAL
    assert relmax(t.eval(g["dx_0_1"], g["points"]), g["eval_dx0"]) <= TOL_EVAL
    assert t.launch_count > 0
    t.close()
    tf.close()


@pytest.mark.parametrize("name", golden_cases("chebyshev"))
@pytest.mark.parametrize("generic", [0, 1])
def test_golden_case_chebyshev(T, name, generic):
    """basis="chebyshev": V = L U, fnt = lower sweep (inv(L) mat-vec or L solve) + upper sweep (inv(U) mat-vec
    or U back-substitution, transform.py:491-548), ifnt = upper + lower mat-vec (:596-624), eval by the
    T_k recurrence (utils.py:165-195).  Bit-exact except eval (tree-shaped sum) like the Newton cases."""
    g = load_case(name)
    t = make(T, g, exact_ifnt=True)
    t.set_option("force_generic", generic)
    m = int(g["m"])
    assert t.basis == "chebyshev"
    assert np.array_equal(t.multi_index_set, g["A"].astype(np.int64)) and np.array_equal(t.nodes, g["x"])
    got = t.fnt(g["values"])
    nbits = int(np.sum(got != g["fnt"]))
    assert nbits == 0, f"{nbits} coefficients differ, rel {relmax(got, g['fnt']):.2e}"
    assert np.array_equal(t.fnt(g["values"][0]), g["fnt"][0])
    assert np.array_equal(t.fnt(g["rnd"]), g["fnt_rnd"])
    assert np.array_equal(t.ifnt(g["fnt"]), g["ifnt"])
    assert np.array_equal(t.ifnt(g["rnd"]), g["ifnt_rnd"])
    tf = make(T, g)
    tf.set_option("force_generic", generic)
    assert relmax(tf.ifnt(g["fnt"]), g["ifnt"]) <= TOL_IFNT_FMA
    for key in g.files:
        if key.startswith("dx_"):
            _, i, k = key.split("_")
            out = t.dx(g["fnt"][0], int(i), int(k))
            assert np.array_equal(out, g[key]), key
        elif key.startswith("dxT_"):
            _, i, k = key.split("_")
            assert np.array_equal(t.dxT(g["rnd"], int(i), int(k)), g[key]), key
    assert relmax(t.eval(g["fnt"][0], g["points"]), g["eval"]) <= TOL_EVAL
    assert relmax(t.eval(g["dx_0_1"], g["points"]), g["eval_dx0"]) <= TOL_EVAL
    # device-tensor path runs the same two sweeps back to back on the caller's stream
    import torch

    c = t.fnt(torch.as_tensor(g["values"], device="cuda"))
    assert np.array_equal(c.cpu().numpy(), g["fnt"])
    assert np.array_equal(t.ifnt(c).cpu().numpy(), g["ifnt"])
    t.close()
    tf.close()


def test_chebyshev_vs_oracle_larger(T):
    """Seeded inputs at a size no fixture holds: Transform(3, 24, basis="chebyshev"), both fnt variants."""
    from oracle import oracle as O

    rng = np.random.default_rng(11)
    for pre in (True, False):
        t = T(3, 24, basis="chebyshev", precomputation=pre, report=False, exact_ifnt=True)
        o = O.OracleTransform(3, 24, 2.0, precomputation=pre, basis="chebyshev")
        v = rng.standard_normal((3, len(t)))
        c = t.fnt(v)
        assert np.array_equal(c, np.stack([o.fnt(r) for r in v]))
        assert np.array_equal(t.ifnt(c), np.stack([o.ifnt(r) for r in c]))
        pts = rng.uniform(-1, 1, (500, 3))
        assert relmax(t.eval(c[0], pts), o.eval(c[0], pts)) <= TOL_EVAL
        t.close()


@pytest.mark.parametrize("whole", [1, -2])
@pytest.mark.parametrize("name", ["d3n7_cheb", "d2n8_cheb", "d2n6_inf_cheb", "d3n7_cheb_solve", "d3n6_inf", "c1_d2n10", "d3n12_p1", "d3n9_p05"])
def test_whole_function_kernels_more_shapes(T, name, whole):
    """The one-pass whole-function kernel (k_whole, two thread groups per SM) on upper-triangular
    sweeps (Chebyshev), p = inf / 1 / 0.5 sets and m = 2; batched so that every group loops over several functions."""
    g = load_case(name)
    t = make(T, g, exact_ifnt=True)
    if whole == -2:  # the slab kernel (coordinates 0 and 1 in one pass), forced on
        t.set_option("use_whole", 0)
        t.set_option("use_slab", 2)
    else:
        t.set_option("use_whole", whole)
    reps = 40
    got = t.fnt(np.tile(g["values"], (reps, 1)))
    assert np.array_equal(got, np.tile(g["fnt"], (reps, 1)))
    back = t.ifnt(np.tile(g["fnt"], (reps, 1)))
    assert np.array_equal(back, np.tile(g["ifnt"], (reps, 1)))
    assert np.array_equal(t.fnt(g["rnd"]), g["fnt_rnd"])
    t.close()


@pytest.mark.parametrize("name", ["c2_d3n20_leja", "d4n8", "d5n6", "d2n7_inf", "d1n12", "d3n12_solve"])
def test_generic_kernels_match_too(T, name):
    """force_generic routes every op through the unbounded-n fallback kernels."""
    g = load_case(name)
    t = make(T, g, exact_ifnt=True)
    t.set_option("force_generic", 1)
    assert np.array_equal(t.fnt(g["values"]), g["fnt"])
    assert np.array_equal(t.ifnt(g["fnt"]), g["ifnt"])
    m = int(g["m"])
    assert np.array_equal(t.dx(g["fnt"][0], m - 1, 1), g[f"dx_{m-1}_1"])
    assert np.array_equal(t.dxT(g["rnd"], 0, 1), g["dxT_0_1"])
    assert relmax(t.eval(g["fnt"][0], g["points"]), g["eval"]) <= TOL_EVAL
    t.close()


def test_device_tensor_path_and_inplace(T):
    import torch

    g = load_case("c2_d3n20_leja")
    t = make(T, g, exact_ifnt=True)
    v = torch.as_tensor(g["values"], device="cuda")
    c = t.fnt(v)
    assert c.is_cuda and np.array_equal(c.cpu().numpy(), g["fnt"])
    r = t.ifnt(c)
    assert np.array_equal(r.cpu().numpy(), g["ifnt"])
    d = t.dx(c[0], 2, 1)
    assert np.array_equal(d.cpu().numpy(), g["dx_2_1"])
    e = t.eval(c[0], torch.as_tensor(g["points"], device="cuda"))
    assert relmax(e.cpu().numpy(), g["eval"]) <= TOL_EVAL
    # in-place through the raw C ABI (in == out), with the colex permutation active
    from lpfun_b200 import _native as nat

    buf = v.clone()
    stream = torch.cuda.current_stream().cuda_stream
    nat.check(nat.lib().lpfnt_fnt(t._plan, buf.data_ptr(), buf.data_ptr(), buf.shape[0], 1, stream))
    assert np.array_equal(buf.cpu().numpy(), g["fnt"])
    nat.check(nat.lib().lpfnt_ifnt(t._plan, buf.data_ptr(), buf.data_ptr(), buf.shape[0], 0, 1, stream))
    assert np.array_equal(buf.cpu().numpy(), g["ifnt"])
    t.close()


def test_operator_level_entry_points(T):
    """lpfnt_itransform / lpfnt_transform / lpfnt_dtransform with caller-supplied packed matrices."""
    from lpfun_b200 import _native as nat

    g = load_case("d4n8")
    t = make(T, g)
    L = nat.lib()
    x = np.ascontiguousarray(g["values"][0])
    out = np.empty_like(x)
    inv = np.ascontiguousarray(g["inv_Vx_lt"])
    nat.check(L.lpfnt_itransform(t._plan, inv.ctypes.data, nat.LPFNT_LOWER, nat.ARITH_EXACT, x.ctypes.data, out.ctypes.data, 1, 0, 0, 0, None))
    assert np.array_equal(out, g["fnt"][0])  # this case has colex_order=False
    gs = load_case("d4n8_solve")
    V = np.ascontiguousarray(gs["Vx_lt"])
    nat.check(L.lpfnt_transform(t._plan, V.ctypes.data, nat.LPFNT_LOWER, x.ctypes.data, out.ctypes.data, 1, 0, 0, 0, None))
    assert np.array_equal(out, gs["fnt"][0])
    D = np.ascontiguousarray(g["Dx_lt1"])
    c = np.ascontiguousarray(g["fnt"][0])
    nat.check(L.lpfnt_dtransform(t._plan, D.ctypes.data, nat.LPFNT_UPPER, nat.ARITH_EXACT, 3, c.ctypes.data, out.ctypes.data, 1, 0, None))
    assert np.array_equal(out, g["dx_3_2"])
    with pytest.raises(ValueError):
        nat.check(L.lpfnt_dtransform(t._plan, D.ctypes.data, nat.LPFNT_UPPER, 0, 4, c.ctypes.data, out.ctypes.data, 1, 0, None))
    t.close()


def test_error_behaviour(T):
    with pytest.raises(ValueError):
        T(0, 4, report=False)
    with pytest.raises(ValueError):
        T(2, 4, 3.0, report=False)
    with pytest.raises(ValueError):
        T(2, 4, basis="legendre", report=False)
    with pytest.raises(ValueError):
        T(3, 30, threshold=100, report=False)
    with pytest.raises(ValueError):
        T(2, 4, nodes=lambda k: np.zeros(k), report=False)
    t = T(2, 4, report=False)
    z = np.zeros(len(t))
    with pytest.raises(ValueError):
        t.dx(z, 2)
    with pytest.raises(ValueError):
        t.dx(z, 0, 4)
    with pytest.raises(ValueError):
        t.fnt(np.zeros(len(t) + 5))  # the reference corrupts the heap here
    with pytest.raises(ValueError):
        t.eval(z, np.zeros((3, 5)))
    assert t.eval(z, np.zeros((0, 2))).shape == (0,)
    t.close()


# ---------------------------------------------------------------------------------------
# the reference's own property tests (test.py:45-151) over its own parameter grid (test.py:8-13): both bases, both
# precomputation modes
# ---------------------------------------------------------------------------------------
GRID = [(m, p, basis, pr) for m in (1, 2, 3, 4, 5, 6) for p in (1.0, 2.0, np.inf) for basis in ("newton", "chebyshev") for pr in (True, False)]


@pytest.mark.parametrize("m, p, basis, pr", GRID)
def test_reference_property_suite(T, m, p, basis, pr):
    rng = np.random.default_rng(100 * m + int(pr))
    for n in (4, 6, 8) if m < 6 else (4, 5):
        if np.isinf(p) and (n + 1) ** m > 300000:
            continue
        t = T(m, n, p, basis=basis, precomputation=pr, precompilation=False, colex_order=False, report=False)
        N = len(t)
        # test_fnt_ifnt
        f = rng.random(N)
        assert np.linalg.norm(t.ifnt(t.fnt(f)) - f) < 1e-8
        # test_dx
        vals = np.sum(t.grid ** (n - 1), axis=1)
        c = t.fnt(vals)
        for k in (1, 2, 3):
            fac = np.prod([n - 1 - q for q in range(k)])
            for i in range(m):
                want = fac * t.grid[:, i] ** (n - 1 - k)
                assert np.linalg.norm(t.ifnt(t.dx(c, i, k)) - want) < 1e-6
        # test_dxT (adjoint identity)
        for k in (1, 2, 3):
            for i in range(m):
                x, y = rng.standard_normal(N), rng.standard_normal(N)
                assert abs(np.dot(t.dx(x, i, k), y) - np.dot(x, t.dxT(y, i, k))) < 1e-6
        t.close()
        # test_eval (degree n+1, cubic)
        t = T(m, n + 1, p, basis=basis, precomputation=pr, precompilation=False, colex_order=False, report=False)
        c = t.fnt(np.sum(t.grid ** 3, axis=1))
        pts = rng.random((10, m))
        assert np.max(np.abs(t.eval(c, pts) - np.sum(pts ** 3, axis=1))) < 1e-6
        t.close()


def test_tutorial_short_config1(T):
    """BASELINE config 1: Transform(2, 10), sin(x)cos(y): fnt -> dx -> eval / ifnt."""
    t = T(spatial_dimension=2, polynomial_degree=10, report=False)
    g = load_case("c1_d2n10")
    v = np.sin(t.grid[:, 0]) * np.cos(t.grid[:, 1])
    assert np.array_equal(v, g["values"][0])
    c = t.fnt(v)
    d = t.dx(c, i=0, k=1)
    assert np.array_equal(c, g["fnt"][0]) and np.array_equal(d, g["dx_0_1"])
    rec = t.ifnt(d)
    assert np.max(np.abs(rec - np.cos(t.grid[:, 0]) * np.cos(t.grid[:, 1]))) < 1e-6
    pts = g["points"]
    assert np.max(np.abs(t.eval(d, pts) - np.cos(pts[:, 0]) * np.cos(pts[:, 1]))) < 1e-6
    t.close()


def test_c3_d6n20_sampled(T):
    """BASELINE config 3 at FULL size (N = 6 974 412): sampled outputs of the reference, plus a
    size-independent property (round trip)."""
    import hashlib

    g = np.load(os.path.join(GOLDEN, "c3_d6n20.npz"))
    t = T(6, 20, report=False, precompilation=False, exact_ifnt=True)
    assert len(t) == int(g["N"])
    gr = t.grid
    w = np.arange(1, 7, dtype=np.float64) / 6.0
    vals = np.exp(-np.sum(gr * gr, axis=1)) * np.sin(gr @ w)
    idx = g["idx"]
    assert np.array_equal(vals[idx], g["values_s"])
    c = t.fnt(vals)
    assert np.array_equal(c[idx], g["fnt_s"])
    assert hashlib.sha256(c.tobytes()).hexdigest() == str(g["fnt_sha"])
    r = t.ifnt(c)
    assert hashlib.sha256(r.tobytes()).hexdigest() == str(g["ifnt_sha"])
    assert np.max(np.abs(r - vals)) == float(g["roundtrip_err"])
    t.close()


def test_c4_batched_vs_oracle(T):
    """BASELINE config 4 shape (d=3, n=30), a batch of 64 seeded functions: CUDA vs the CPU oracle."""
    from oracle import oracle as O

    t = T(3, 30, report=False, precompilation=False, exact_ifnt=True)
    o = O.OracleTransform(3, 30, 2.0)
    assert np.array_equal(o.P, t.colex_order) and np.array_equal(o.grid, t.grid)
    rng = np.random.default_rng(1234)
    B = 64
    a = rng.uniform(0.5, 1.0, size=(B, 4))
    w = rng.uniform(-3, 3, size=(B, 4, 3))
    ph = rng.uniform(0, 2 * np.pi, size=(B, 4))
    vals = np.stack([np.sum(a[b][None, :] * np.sin(t.grid @ w[b].T + ph[b][None, :]), axis=1) for b in range(B)])
    c = t.fnt(vals)
    r = t.ifnt(c)
    for b in range(0, B, 7):
        cb = o.fnt(vals[b])
        assert np.array_equal(c[b], cb)
        assert np.array_equal(r[b], o.ifnt(cb))
    # linearity of the batched path (size-independent property)
    s = t.ifnt(c[0] + 2.0 * c[1])
    assert relmax(s, r[0] + 2.0 * r[1]) < 1e-12
    t.close()


def test_c5_dx_eval_vs_oracle(T):
    """BASELINE config 5 shape (d=4, n=20): dx along every coordinate and eval at 4096 random points."""
    from oracle import oracle as O

    g = load_case("c5_d4n20")
    t = make(T, g)
    c = g["fnt"][0]
    o = O.Sweeper(g["A"].astype(np.int64))
    rng = np.random.default_rng(1234)
    pts = rng.random((4096, 4)) * 2 - 1
    for i in range(4):
        d = t.dx(c, i, 1)
        assert np.array_equal(d, g[f"dx_{i}_1"])
    want = o.eval(g["dx_0_1"], g["x"], pts, 20)
    for split in (0, 1, -1):  # block-staged kernel / split path (subtrees over CTAs + basis-form combination) / automatic
        t.set_option("eval_split", split)
        assert relmax(t.eval(g["dx_0_1"], pts), want) <= TOL_EVAL, split
    t.close()


@pytest.mark.parametrize("m,n,p", [(2, 10, 2.0), (3, 12, 1.0), (3, 30, 2.0), (4, 9, np.inf), (5, 6, 2.0), (6, 8, 2.0), (2, 31, 2.0)])
@pytest.mark.parametrize("P", [1, 10, 300])
def test_eval_few_points_split_path(T, m, n, p, P):
    """Few points: the index trie is cut into subtrees that are evaluated side by side and combined in basis form
    (k_eval_sub + k_eval_combine) -- against the oracle's reference-order sum and against the one-thread-per-point kernels."""
    from oracle import oracle as O

    rng = np.random.default_rng(31 * m + n + P)
    t = T(m, n, lp_degree=p, report=False)
    o = O.OracleTransform(m, n, p)
    c = rng.standard_normal(len(t)) / (1.0 + np.arange(len(t))) ** 0.5
    pts = rng.uniform(-1, 1, (P, m))
    want = o.eval(c, pts)
    t.set_option("eval_split", 1)
    l0 = t.launch_count
    got = t.eval(c, pts)
    assert (t.launch_count - l0) % 2 == 0 and t.launch_count > l0  # the split path really ran (two launches per chunk of points)
    tol = 1e-12 if n < 25 else 1e-10
    assert relmax(got, want) <= tol
    import torch

    dev = t.eval(torch.as_tensor(c, device="cuda"), torch.as_tensor(pts, device="cuda"))  # device tensors: same kernels
    assert np.array_equal(dev.cpu().numpy(), got)
    t.set_option("eval_split", 0)
    assert relmax(t.eval(c, pts), got) <= tol
    t.close()


@pytest.mark.parametrize("opts", [
    {"use_whole": 0},
    {"use_whole": 1},
    {"use_whole": 1, "whole_threads": 384},
    {"use_whole": 1, "whole_threads": 384, "whole_order": 0},
    {"use_whole": 1, "whole_order": 63},
    {"use_whole": 1, "whole_threads": 384, "whole_fibres": 2},
    {"use_whole": 1, "whole_threads": 384, "whole_fibres": 2, "whole_order": 0},
    {"use_whole": 1, "whole_threads_exact": 384, "whole_fibres_exact": 2, "whole_threads_fma": 768, "whole_fibres_fma": 1},
    {"use_whole": 1, "whole_split": 1},
    {"use_whole": 1, "whole_split": 1, "whole_split_c": 7},
    {"use_whole": 1, "whole_split": 1, "whole_split_order": 64},
    {"use_whole": 1, "whole_threads": 768, "whole_fibres": 1},
    {"use_whole": 0, "use_slab": 2},
    {"use_whole": 0, "use_slab": 0},
    {"use_whole": 0, "use_cta_cap": 0},
    {},
])
@pytest.mark.parametrize("name", ["c2_d3n20_leja", "d4n8", "d5n6", "d6n4", "d2n7_inf", "d3n12_solve", "c4_d3n30"])
def test_alternative_kernel_families(T, name, opts):
    """The selectable kernel paths (per-coordinate kernels, slab kernel, whole-function kernel
    with 384 or 768 threads per CTA, one or two fibres per thread, other item orders) must all reproduce the reference bit for bit."""
    g = load_case(name)
    t = make(T, g, exact_ifnt=True)
    for k, v in opts.items():
        t.set_option(k, v)
    assert np.array_equal(t.fnt(g["values"]), g["fnt"])
    assert np.array_equal(t.ifnt(g["fnt"]), g["ifnt"])
    assert np.array_equal(t.fnt(np.tile(g["rnd"], (5, 1)))[3], g["fnt_rnd"])
    m = int(g["m"])
    assert np.array_equal(t.dx(g["fnt"][0], m - 1, 1), g[f"dx_{m-1}_1"])
    assert np.array_equal(t.dxT(g["rnd"], m - 1, 1), g[f"dxT_{m-1}_1"])
    t.close()


def test_prolong_restrict_through_embed(T):
    """Device coefficient transfer through Transform.embed (transform.py:850-899): the prolonged coefficients
    describe the same polynomial, restrict undoes prolong, host and device paths agree."""
    import torch

    rng = np.random.default_rng(5)
    for m, n_c, n_f, p_c, p_f in ((2, 6, 9, 2.0, 2.0), (3, 5, 8, 1.0, 2.0), (4, 3, 5, 2.0, np.inf)):
        from lpfun_b200.utils import leja_nodes

        tc = T(m, n_c, lp_degree=p_c, report=False, nodes=leja_nodes)
        tf = T(m, n_f, lp_degree=p_f, report=False, nodes=leja_nodes)
        phi = tc.embed(tf)
        c = rng.standard_normal((3, len(tc)))
        f = tc.prolong(c, tf)
        want = np.zeros((3, len(tf)))
        want[:, phi] = c
        assert np.array_equal(f, want)
        assert np.array_equal(tc.restrict(f, tf), c)
        fd = tc.prolong(torch.as_tensor(c[0], device="cuda"), tf)
        assert fd.is_cuda and np.array_equal(fd.cpu().numpy(), want[0])
        pts = rng.uniform(-1, 1, (64, m))
        assert relmax(tf.eval(f[0], pts), tc.eval(c[0], pts)) <= 1e-11
        with pytest.raises(ValueError):
            tf.prolong(f, tc)
        tc.close(); tf.close()


def _run_sharded(world):
    import subprocess
    import sys

    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_sharded_gpu.py")
    if world == 1:
        cmd = [sys.executable, script]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", "29541", script]
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _run_push(world):
    import subprocess
    import sys

    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_push_gpu.py")
    if world == 1:
        cmd = [sys.executable, script]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", "29543", script]
    env = dict(os.environ)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "PUSH_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_push_gather_one_gpu():
    """The whole-function kernel storing its result into the symmetric buffer (a world of one: the own mapping only)."""
    _run_push(1)


def test_push_gather_two_gpus():
    """Row blocks of a sharded batch gathered by the last sweep's stores into every rank's buffer (multimem.st / peer stores)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    _run_push(2)


def test_sharded_single_transform_one_gpu():
    """Slab sweeps -> all-to-all -> column sweep on a process group of one (NCCL): bit-identical to the plain transform."""
    _run_sharded(1)


def test_sharded_single_transform_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    _run_sharded(2)


@pytest.mark.parametrize("m,n,p", [(2, 31, 2.0), (3, 31, 2.0), (2, 32, 2.0), (2, 40, 1.0), (3, 33, 2.0), (1, 31, 2.0), (1, 45, 2.0),
                                   (3, 30, 1.0), (3, 29, 2.0), (2, 1, 2.0), (3, 1, np.inf), (5, 2, 2.0)])
def test_size_boundaries_vs_oracle(T, m, n, p):
    """Capacity edges: n + 1 = 32 is the last register-resident size, above it the generic kernels take over;
    tiny sets; odd N / N_1 (the whole-function kernel's eligibility rules)."""
    from oracle import oracle as O

    rng = np.random.default_rng(100 * m + n)
    t = T(m, n, lp_degree=p, report=False, exact_ifnt=True)
    o = O.OracleTransform(m, n, p)
    assert np.array_equal(t.multi_index_set, o.A) and len(t) == o.N
    v = rng.standard_normal((3, len(t)))
    c = t.fnt(v)
    assert np.array_equal(c, np.stack([o.fnt(r) for r in v]))
    assert np.array_equal(t.ifnt(c), np.stack([o.ifnt(r) for r in c]))
    if n >= 1:
        assert np.array_equal(t.dx(c[0], m - 1, 1), o.dx(c[0], m - 1, 1))  # p = inf: folded descending, like the reference
        assert np.array_equal(t.dxT(v[0], 0, 1), o.dxT(v[0], 0, 1))
    pts = rng.uniform(-1, 1, (33, m))
    # eval on random (non-smooth) coefficients at n >= 25 mixes terms of very different size, and the reference's own flat
    # left-to-right sum loses digits there: judge both against an extended-precision truth (x87 80-bit long double, every
    # product and sum in long double) -- the GPU's nested Horner form must be at least as accurate as the reference order
    truth = eval_truth_longdouble(c[0], o.x, pts, o.A)
    scale = float(np.max(np.abs(truth)))
    got, ref = t.eval(c[0], pts), o.eval(c[0], pts)
    err_gpu = float(np.max(np.abs(got.astype(np.longdouble) - truth)))
    err_ref = float(np.max(np.abs(ref.astype(np.longdouble) - truth)))
    assert err_gpu <= max(err_ref, 1e-12 * scale), (err_gpu / scale, err_ref / scale)
    if n < 25:
        assert relmax(got, ref) <= 1e-12
    t.close()


def eval_truth_longdouble(c, nodes, pts, A):
    """newton2point (src/lpfun/utils.py:132-162) in np.longdouble: basis tables by the same recurrence, flat sum."""
    ld = np.longdouble
    P, m = pts.shape
    n1 = len(nodes)
    basis = np.ones((P, m, n1), dtype=ld)
    for q in range(1, n1):
        basis[:, :, q] = basis[:, :, q - 1] * (pts.astype(ld) - ld(nodes[q - 1]))
    out = np.zeros(P, dtype=ld)
    cl = c.astype(ld)
    for b0 in range(0, len(A), 4096):
        blk = A[b0:b0 + 4096]
        prod = np.ones((P, len(blk)), dtype=ld)
        for j in range(m):
            prod *= basis[:, j, :][:, blk[:, j]]
        out += prod @ cl[b0:b0 + 4096]
    return out


@pytest.mark.parametrize("chunk_bytes,threads", [(1 << 20, 0), (1 << 20, 1), (3 << 20, 3), (0, 0)])
def test_pageable_host_arrays_are_staged(T, chunk_bytes, threads):
    """Plain (pageable) NumPy arrays of >= 4 MB go through the plan's pinned staging buffers, copied by host threads while
    the other lane's chunk is on the GPU; the result must equal the pinned-buffer path bit for bit, for every chunking."""
    from lpfun_b200 import _native as nat

    g = load_case("c4_d3n30")
    t = make(T, g, exact_ifnt=True)
    N = len(t)
    B = 67  # 8.2 MB: odd number of functions, last chunk partial
    rng = np.random.default_rng(3)
    v = rng.standard_normal((B, N))
    hv = nat.pinned_empty((B, N)); hc = nat.pinned_empty((B, N)); hr = nat.pinned_empty((B, N))
    hv[:] = v
    t.fnt(hv, out=hc); t.ifnt(hc, out=hr)
    t.set_option("host_chunk_bytes", chunk_bytes)
    t.set_option("host_threads", threads)
    c = t.fnt(v)           # pageable in, fresh pageable out
    assert np.array_equal(c, hc)
    r = t.ifnt(c)
    assert np.array_equal(r, hr)
    out = np.empty_like(v)
    t.fnt(hv, out=out)     # pinned in, pageable out
    assert np.array_equal(out, hc)
    t.ifnt(c, out=hr)      # pageable in, pinned out
    assert np.array_equal(hr, r)
    for a in (hv, hc, hr):
        nat.pinned_free(a)
    t.close()


def test_batch_beyond_the_grid_limit(T):
    """More functions than gridDim.y allows in one launch (65 535): the launchers loop."""
    from oracle import oracle as O

    t = T(2, 4, report=False, exact_ifnt=True)
    o = O.OracleTransform(2, 4, 2.0)
    rng = np.random.default_rng(9)
    B = 70001
    v = rng.standard_normal((B, len(t)))
    c = t.fnt(v)
    for b in (0, 65534, 65535, 65536, B - 1):
        assert np.array_equal(c[b], o.fnt(v[b]))
    r = t.ifnt(c)
    for b in (0, 65535, B - 1):
        assert np.array_equal(r[b], o.ifnt(c[b]))
    t.close()


@pytest.mark.parametrize("m,n,p", [(1, 9, 2.0), (2, 12, 2.0), (3, 10, 1.0), (3, 23, 2.0), (4, 8, np.inf), (5, 6, 2.0), (3, 30, 2.0)])
def test_chebyshev_eval_fast_kernel_vs_oracle(T, m, n, p):
    """Block-staged Chebyshev evaluation (T_k table for coordinate 0, Clenshaw over the trie for the others) against the
    reference-order flat sum of the oracle and against the generic fallback kernel."""
    from oracle import oracle as O

    rng = np.random.default_rng(17 * m + n)
    t = T(m, n, lp_degree=p, basis="chebyshev", report=False)
    o = O.OracleTransform(m, n, p, basis="chebyshev")
    c = rng.standard_normal(len(t)) / (1.0 + np.arange(len(t)))
    pts = rng.uniform(-1, 1, (777, m))
    want = o.eval(c, pts)
    got = t.eval(c, pts)
    assert relmax(got, want) <= TOL_EVAL
    t.set_option("force_generic", 1)
    assert relmax(t.eval(c, pts), want) <= TOL_EVAL
    t.close()


def test_device_calls_are_cuda_graph_capturable(T):
    """Device-pointer calls only enqueue on the caller's stream and allocate nothing once the plan is warm: a loop of dx /
    fnt / ifnt calls can be captured in a CUDA graph (launch-bound loops: 6.8 instead of 23 us per dx at d=4, n=20) and
    the replay reproduces the eager results bit for bit."""
    import torch

    t = T(4, 9, report=False)
    rng = np.random.default_rng(5)
    c = torch.as_tensor(rng.standard_normal(len(t)), device="cuda")
    d = [torch.empty_like(c) for _ in range(4)]
    v = torch.empty_like(c); r = torch.empty_like(c)
    for i in range(4):
        t.dx(c, i, 1, out=d[i])
    t.ifnt(c, out=v); t.fnt(v, out=r)
    torch.cuda.synchronize()
    want = [x.clone() for x in d] + [v.clone(), r.clone()]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(4):
            t.dx(c, i, 1, out=d[i])
        t.ifnt(c, out=v); t.fnt(v, out=r)
    for x in d + [v, r]:
        x.zero_()
    g.replay()
    torch.cuda.synchronize()
    for got, w in zip(d + [v, r], want):
        assert torch.equal(got, w)
    c.mul_(2.0)  # the graph reads the buffers it was captured on
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(d[0], 2.0 * want[0])
    t.close()


@pytest.mark.parametrize("basis", ["newton", "chebyshev"])
@pytest.mark.parametrize("m,n,p", [(7, 5, 2.0), (9, 3, 2.0), (12, 2, 1.0), (8, 4, np.inf), (3, 31, 2.0), (1, 31, 2.0)])
def test_eval_tree_kernel_many_levels_and_full_capacity(T, m, n, p, basis):
    """k_eval_tree at the ends of its template grid: 5..12 coordinates (the instantiations with 8 and 12 fold levels), the
    full 32-entry node table, a single coordinate; Newton and Chebyshev; the one-thread-per-point path forced (eval_split 0);
    a point count that is not a multiple of the CTA's 256 points."""
    from oracle import oracle as O

    rng = np.random.default_rng(101 * m + n)
    t = T(m, n, lp_degree=p, basis=basis, report=False)
    o = O.OracleTransform(m, n, p, basis=basis)
    c = rng.standard_normal(len(t)) / (1.0 + np.arange(len(t))) ** 0.5
    pts = rng.uniform(-1, 1, (517, m))
    want = o.eval(c, pts)
    t.set_option("eval_split", 0)
    got = t.eval(c, pts)
    tol = TOL_EVAL if n < 25 else 1e-10
    assert relmax(got, want) <= tol
    if len(t) * m <= (1 << 20):  # (larger sets do not keep the multi-index array the fallback kernel walks)
        t.set_option("force_generic", 1)
        assert relmax(t.eval(c, pts), want) <= tol
    t.close()


@pytest.mark.parametrize("n,B", [(25, 1), (25, 16), (28, 9), (29, 300), (12, 7)])
def test_whole_function_kernel_odd_sizes(T, n, B):
    """k_whole with an odd number of coefficients: every second function starts 8 bytes off the 16-byte grid of the
    vectorised copies; one or two CTAs per SM, batches smaller and larger than the grid."""
    from oracle import oracle as O

    t = T(3, n, report=False, exact_ifnt=True)
    assert len(t) % 2 == 1
    t.set_option("use_whole", 1)
    t.set_option("whole_threads", 384 if (n + B) % 2 else 768)
    o = O.OracleTransform(3, n, 2.0)
    rng = np.random.default_rng(n + B)
    v = rng.standard_normal((B, len(t)))
    c = t.fnt(v)
    r = t.ifnt(c)
    for b in sorted({0, 1, B // 2, B - 2, B - 1} & set(range(B))):
        assert np.array_equal(c[b], o.fnt(v[b])), b
        assert np.array_equal(r[b], o.ifnt(c[b])), b
    t.close()

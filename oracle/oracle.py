"""ctypes front-end of the CPU oracle (oracle/lpfnt_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by lpfun_b200/.

Parity status: PINNED.  tests/test_oracle_golden.py compares this oracle
bit-for-bit with outputs of the unmodified reference stored in tests/golden/
(tables for the reference's whole test grid, fnt/ifnt/dx/dxT/eval vectors for
26 Transforms incl. the five BASELINE configs).

``OracleTransform`` mirrors the reference ``Transform`` front-end
(src/lpfun/transform.py:95-928) for both bases on top of the C sweeps.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liblpfnt_oracle.so")
_lib = None

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "lpfnt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_lp_set.restype = C.c_int64
        L.orc_lp_set.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int64]
        L.orc_lp_tube.restype = C.c_int64
        L.orc_lp_tube.argtypes = [_i64p, C.c_int64, C.c_int, C.c_void_p]
        L.orc_neighbors.argtypes = [_i64p, C.c_int64, C.c_int, C.c_int, _i32p, _i32p]
        L.orc_sweep_lower_matvec.argtypes = [_f64p, _f64p, _f64p, _i64p, C.c_int64, C.c_int, C.c_int, _i32p]
        L.orc_sweep_upper_matvec.argtypes = [_f64p, C.c_int, _f64p, _f64p, _i64p, C.c_int64, C.c_int, C.c_int, _i32p, C.c_int]
        L.orc_sweep_lower_solve.argtypes = [_f64p, _f64p, _f64p, _i64p, C.c_int64, C.c_int, C.c_int, _i32p]
        L.orc_sweep_upper_solve.argtypes = [_f64p, C.c_int, _f64p, _f64p, _i64p, C.c_int64, C.c_int, C.c_int, _i32p]
        L.orc_chebyshev2point.argtypes = [_f64p, _f64p, C.c_int64, _i64p, C.c_int64, C.c_int, C.c_int, _f64p]
        L.orc_newton2point.argtypes = [_f64p, _f64p, _f64p, C.c_int64, _i64p, C.c_int64, C.c_int, C.c_int, _f64p]
        L.orc_roundtrip_batch.argtypes = [_f64p, _f64p, _f64p, _f64p, _f64p, _i64p, C.c_int64, C.c_int, _i32p, C.c_void_p, C.c_int64]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def use_all_cores() -> int:
    """Use every host core (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(n)
    return num_threads()


def lp_set(m: int, n: int, p: float) -> np.ndarray:
    L = lib()
    N = L.orc_lp_set(int(m), int(n), float(p), None, 0)
    A = np.zeros((N, m), dtype=np.int64)
    got = L.orc_lp_set(int(m), int(n), float(p), A.ctypes.data, N)
    assert got == N
    return A


def lp_tube(A: np.ndarray) -> np.ndarray:
    A = np.ascontiguousarray(A, dtype=np.int64)
    L = lib()
    n1 = L.orc_lp_tube(A, len(A), A.shape[1], None)
    T = np.zeros(n1, dtype=np.int64)
    L.orc_lp_tube(A, len(A), A.shape[1], T.ctypes.data)
    return T


class Sweeper:
    """Neighbour tables of one multi-index set + the three sweep kinds."""

    def __init__(self, A: np.ndarray):
        self.A = np.ascontiguousarray(A, dtype=np.int64)
        self.N, self.m = self.A.shape
        self._nb = {}

    def neighbors(self, h: int):
        if h not in self._nb:
            prev = np.empty(self.N, dtype=np.int32)
            nxt = np.empty(self.N, dtype=np.int32)
            lib().orc_neighbors(self.A, self.N, self.m, h, prev, nxt)
            self._nb[h] = (prev, nxt)
        return self._nb[h]

    def lower_matvec(self, Lp, x, dims=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        Lp = np.ascontiguousarray(Lp, dtype=np.float64)
        for h in (range(self.m) if dims is None else dims):
            out = np.empty_like(x)
            lib().orc_sweep_lower_matvec(Lp, x, out, self.A, self.N, self.m, h, self.neighbors(h)[0])
            x = out
        return x

    def lower_solve(self, Lp, x, dims=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        Lp = np.ascontiguousarray(Lp, dtype=np.float64)
        for h in (range(self.m) if dims is None else dims):
            out = np.empty_like(x)
            lib().orc_sweep_lower_solve(Lp, x, out, self.A, self.N, self.m, h, self.neighbors(h)[1])
            x = out
        return x

    def upper_matvec(self, Up, n1, x, dims=None, descending=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        Up = np.ascontiguousarray(Up, dtype=np.float64)
        for h in (range(self.m) if dims is None else dims):
            out = np.empty_like(x)
            lib().orc_sweep_upper_matvec(Up, int(n1), x, out, self.A, self.N, self.m, h,
                                         self.neighbors(h)[1], int(bool(descending)))
            x = out
        return x

    def upper_solve(self, Up, n1, x, dims=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        Up = np.ascontiguousarray(Up, dtype=np.float64)
        for h in (range(self.m) if dims is None else dims):
            out = np.zeros_like(x)
            lib().orc_sweep_upper_solve(Up, int(n1), x, out, self.A, self.N, self.m, h, self.neighbors(h)[1])
            x = out
        return x

    def eval(self, c, nodes, pts, n):
        c = np.ascontiguousarray(c, dtype=np.float64)
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        out = np.empty(len(pts), dtype=np.float64)
        lib().orc_newton2point(c, nodes, pts, len(pts), self.A, self.N, self.m, int(n), out)
        return out

    def eval_chebyshev(self, c, pts, n):
        c = np.ascontiguousarray(c, dtype=np.float64)
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        out = np.empty(len(pts), dtype=np.float64)
        lib().orc_chebyshev2point(c, pts, len(pts), self.A, self.N, self.m, int(n), out)
        return out


class OracleTransform:
    """CPU mirror of the reference Transform (Newton and Chebyshev bases).

    Matrices are taken from the caller (golden fixtures) or built with the host
    numerics in lpfun_b200.utils, which are themselves pinned against the
    reference's matrices by tests/test_host_golden.py.
    """

    def __init__(self, m, n, p=2.0, nodes="cheb2nd", precomputation=True, colex_order=True, mats=None, basis="newton"):
        from lpfun_b200 import utils as U  # host numerics only (no GPU code)

        self.m, self.n, self.p = int(m), int(n), float(p)
        self.A = lp_set(self.m, self.n, self.p)
        self.T = lp_tube(self.A)
        self.N = len(self.A)
        self.sw = Sweeper(self.A)
        x = (U.leja_nodes if nodes == "leja" else U.cheb2nd_nodes)(self.n + 1) if isinstance(nodes, str) else nodes(self.n + 1)
        self.leja_order = U.get_leja_order(x)
        self.x = x[self.leja_order]
        grid_int = self.x[self.A]
        if colex_order:
            self.P = np.lexsort(grid_int.T)
            self.grid = np.empty_like(grid_int)
            self.grid[self.P] = grid_int
        else:
            self.P, self.grid = None, grid_int
        self.basis = basis
        if mats is None:
            cheb = basis == "chebyshev"
            Vx = (U.chebyshev2lagrange if cheb else U.newton2lagrange)(self.x)
            Dx = (U.chebyshev2derivative if cheb else U.newton2derivative)(self.x)
            Dx2 = Dx @ Dx
            Dx3 = Dx @ Dx2
            mats = dict(Dx_ut=[U.pack_upper(D) for D in (Dx, Dx2, Dx3)],
                        DxT_lt=[U.pack_lower(D.T) for D in (Dx, Dx2, Dx3)])
            if U.is_lower_triangular(Vx):
                mats.update(Vx_lt=U.pack_lower(Vx), inv_Vx_lt=U.pack_lower(np.linalg.inv(Vx)) if precomputation else None)
            else:  # src/lpfun/transform.py:292-302
                Lf, Uf = U.get_lu(Vx)
                mats.update(Vx_lt=U.pack_lower(Lf), Vx_ut=U.pack_upper(Uf),
                            inv_Vx_lt=U.pack_lower(np.linalg.inv(Lf)) if precomputation else None,
                            inv_Vx_ut=U.pack_upper(np.linalg.inv(Uf)) if precomputation else None)
        self.mats = mats
        self.precomputation = precomputation

    def __len__(self):
        return self.N

    # src/lpfun/transform.py:430-490
    def fnt(self, values):
        v = np.asarray(values, dtype=np.float64)
        if self.P is not None:
            v = v[self.P]
        n1 = self.n + 1
        if self.mats.get("Vx_ut") is not None:  # LU two-pass, transform.py:491-548
            if self.precomputation:
                return self.sw.upper_matvec(self.mats["inv_Vx_ut"], n1, self.sw.lower_matvec(self.mats["inv_Vx_lt"], v))
            return self.sw.upper_solve(self.mats["Vx_ut"], n1, self.sw.lower_solve(self.mats["Vx_lt"], v))
        if self.precomputation:
            return self.sw.lower_matvec(self.mats["inv_Vx_lt"], v)
        return self.sw.lower_solve(self.mats["Vx_lt"], v)

    # src/lpfun/transform.py:553-632
    def ifnt(self, coeffs):
        c = np.asarray(coeffs, dtype=np.float64)
        if self.mats.get("Vx_ut") is not None:  # transform.py:596-624
            c = self.sw.upper_matvec(self.mats["Vx_ut"], self.n + 1, c)
        y = self.sw.lower_matvec(self.mats["Vx_lt"], c)
        if self.P is not None:
            out = np.empty_like(y)
            out[self.P] = y
            return out
        return y

    def roundtrip_batch(self, values):
        """(coefficients, reconstruction) of a batch (B, N): fnt then ifnt, one OpenMP thread per function
        (bench.py's "threads over functions" CPU figure; Newton basis, precomputation=True)."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.ndim == 2 and v.shape[1] == self.N and self.precomputation and self.mats.get("Vx_ut") is None
        prev = np.ascontiguousarray(np.concatenate([self.sw.neighbors(h)[0] for h in range(self.m)]))
        coef = np.empty_like(v)
        out = np.empty_like(v)
        P = None if self.P is None else np.ascontiguousarray(self.P, dtype=np.int64)
        lib().orc_roundtrip_batch(np.ascontiguousarray(self.mats["inv_Vx_lt"]), np.ascontiguousarray(self.mats["Vx_lt"]), v, coef, out,
                                  self.sw.A, self.N, self.m, prev, None if P is None else P.ctypes.data, v.shape[0])
        return coef, out

    # src/lpfun/transform.py:634-718, src/lpfun/core/molecules.py:222-309
    def dx(self, coeffs, i, k=1):
        desc = np.isinf(self.p) and self.m > 1
        return self.sw.upper_matvec(self.mats["Dx_ut"][k - 1], self.n + 1, coeffs, dims=[i], descending=desc)

    # src/lpfun/transform.py:720-804
    def dxT(self, coeffs, i, k=1):
        return self.sw.lower_matvec(self.mats["DxT_lt"][k - 1], coeffs, dims=[i])

    # src/lpfun/transform.py:806-848
    def eval(self, coeffs, points):
        if self.basis == "chebyshev":
            return self.sw.eval_chebyshev(coeffs, points, self.n)
        return self.sw.eval(coeffs, self.x, points, self.n)

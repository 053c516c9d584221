"""``Transform`` -- drop-in for ``lpfun.Transform`` whose hot path runs on a B200.

Same constructor signature, properties and NumPy in/out contract as the reference class
(src/lpfun/transform.py:95-928).  The constructor is the *plan builder*: it enumerates the
l^p lower set, derives the line / fibre tables, builds the packed 1-D Newton operators on
the host (bit-identical to the reference's) and uploads everything once.  ``fnt``, ``ifnt``,
``dx``, ``dxT`` and ``eval`` then call the hand-written sm_100a kernels of ``liblpfnt.so``
through the C ABI in ``include/lpfnt.h``.  There is no CPU fallback.

Extensions over the reference API
---------------------------------
* a leading batch axis: ``fnt/ifnt/dx/dxT`` accept ``(N,)`` or ``(B, N)`` and return the same shape;
* CUDA ``torch.Tensor`` inputs stay on the device (the result is a CUDA tensor; the work is
  enqueued on torch's current stream);
* ``device=`` selects the GPU; ``exact_ifnt=True`` makes ``ifnt`` bit-identical to the reference
  too (default: same summation order with FMA, <= 1e-15 relative deviation).
* input lengths are validated (the reference silently corrupts memory).
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Callable, Literal

import numpy as np

from . import _native as nat
from . import utils as U
from .core import set as lpset
from .core.utils import apply_permutation

try:  # torch is only an optional carrier of device arrays
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_cuda_tensor(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor) and x.is_cuda


class Transform:
    """Fast Newton Transform on an l^p lower set, computed on one B200.

    Parameters mirror src/lpfun/transform.py:119-131.
    """

    def __init__(
        self,
        spatial_dimension: int,
        polynomial_degree: int,
        lp_degree: float = 2.0,
        nodes: Callable[[int], np.ndarray] = U.cheb2nd_nodes,
        basis: Literal["newton", "chebyshev"] = "newton",
        precomputation: bool = True,
        precompilation: bool = True,
        colex_order: bool = True,
        threshold: int | None = 150_000_000,
        report: bool = True,
        *,
        device: int = 0,
        exact_ifnt: bool = False,
    ):
        t0 = time.time()
        self._m = int(spatial_dimension)
        self._n = int(polynomial_degree)
        self._p = float(lp_degree)
        self._basis = str(basis)
        self._device = int(device)
        self._exact_ifnt = bool(exact_ifnt)
        self._plan = None
        U.classify(self._m, self._n, self._p)
        if basis not in ("newton", "chebyshev"):
            raise ValueError("Invalid choice for basis.")

        # --- index structure (transform.py:207-232) ---
        self._A = lpset.lp_set(self._m, self._n, self._p)
        self._N_0 = len(self._A)
        self._T = lpset.lp_tube(self._A)
        self._cs_T = np.concatenate((np.zeros(1, dtype=np.int64), np.cumsum(self._T)))
        self._N_1 = len(self._T)
        self._e_T = lpset.entropy(self._A)

        # --- threshold (transform.py:234-245) ---
        if threshold is not None and self._N_0 > threshold:
            raise ValueError(
                f"Dimension exceeds threshold: {self._N_0:_} > {threshold:_}. "
                "If this operation should be executed anyways, please set threshold to None."
            )

        # --- nodes, Leja order (transform.py:247-254) ---
        x = np.asarray(nodes(self._n + 1), dtype=np.float64)
        if len(np.unique(x)) != self._n + 1:
            raise ValueError("The provided nodes are not pairwise distinct.")
        self._leja_order = U.get_leja_order(x)
        self._x = np.ascontiguousarray(x[self._leja_order])

        # --- grid + colex permutation (transform.py:256-266) ---
        grid = self._x[self._A]  # grid[i, j] = x[A[i, j]]  (utils.py:231-272)
        if colex_order:
            # np.lexsort(grid.T) of the reference (transform.py:259-261): the nodes are distinct, so sorting the points by their
            # node VALUES (last coordinate most significant) equals sorting one integer key per point built from the value
            # ranks of its indices -- the same permutation in a fifth of the time (4.3 -> 0.85 s at d=6, n=20)
            n1 = self._n + 1
            if float(n1) ** self._m < 2.0 ** 62:
                rank = np.empty(n1, dtype=np.int64)
                rank[np.argsort(self._x, kind="stable")] = np.arange(n1)
                key = np.zeros(self._N_0, dtype=np.int64)
                for j in range(self._m - 1, -1, -1):
                    key = key * n1 + rank[self._A[:, j]]
                self._colex_order = np.argsort(key, kind="stable")
            else:
                self._colex_order = np.lexsort(grid.T)
            self._grid = apply_permutation(self._colex_order, grid, invert=False)
        else:
            self._colex_order = None
            self._grid = grid

        # --- 1-D operators, packed exactly like the reference (transform.py:268-336) ---
        if basis == "newton":
            self._Vx = U.newton2lagrange(self._x)
            self._Dx = U.newton2derivative(self._x)
        else:
            self._Vx = U.chebyshev2lagrange(self._x)
            self._Dx = U.chebyshev2derivative(self._x)
        self._Dx2 = self._Dx @ self._Dx
        self._Dx3 = self._Dx @ self._Dx2
        self._cond_Vx = float(np.linalg.cond(self._Vx))
        self._Vx_ut = self._inv_Vx_ut = None
        if U.is_lower_triangular(self._Vx):
            self._Vx_lt = U.pack_lower(self._Vx)
            self._inv_Vx_lt = U.pack_lower(np.linalg.inv(self._Vx)) if precomputation else None
        else:
            # V = L U without pivoting (transform.py:292-302)
            Lf, Uf = U.get_lu(self._Vx)
            self._Vx_lt, self._Vx_ut = U.pack_lower(Lf), U.pack_upper(Uf)
            self._inv_Vx_lt = U.pack_lower(np.linalg.inv(Lf)) if precomputation else None
            self._inv_Vx_ut = U.pack_upper(np.linalg.inv(Uf)) if precomputation else None
        if not U.is_lower_triangular(self._Dx.T):
            raise NotImplementedError("differentiation matrices that are not upper triangular need the LU path (not implemented)")
        self._Dx_lt = [U.pack_upper(D) for D in (self._Dx, self._Dx2, self._Dx3)]  # reference name; upper packed
        self._DxT_lt = [U.pack_lower(D.T) for D in (self._Dx, self._Dx2, self._Dx3)]
        self._construction_ms = (time.time() - t0) * 1000

        # --- device plan ---
        t1 = time.time()
        self._plan = self._create_plan()
        if precompilation:
            self.warmup()
        self._precompilation_ms = (time.time() - t1) * 1000
        if report:
            print(self)

    # ------------------------------------------------------------------ plan
    def _create_plan(self):
        L = nat.lib()
        dptr = (C.c_void_p * 3)(*[a.ctypes.data for a in self._Dx_lt])
        tptr = (C.c_void_p * 3)(*[a.ctypes.data for a in self._DxT_lt])
        plan = C.c_void_p()
        perm = None if self._colex_order is None else np.ascontiguousarray(self._colex_order, dtype=np.int64)
        rc = L.lpfnt_plan_create(
            self._m, self._n, self._N_0, self._A.ctypes.data, nat.ptr(perm), self._x.ctypes.data,
            self._Vx_lt.ctypes.data, nat.ptr(self._inv_Vx_lt), dptr, tptr, self._device, C.byref(plan),
        )
        nat.check(rc)
        if self._Vx_ut is not None:
            nat.check(L.lpfnt_plan_set_upper_factors(plan, self._Vx_ut.ctypes.data, nat.ptr(self._inv_Vx_ut), int(self._basis == "chebyshev")))
        if np.isinf(self._p) and self._m > 1:
            # the reference differentiates p = inf sets on reversed arrays (molecules.py:290-294): every row of dx is
            # folded from its last term down; reproduced so that dx is bit-identical there too
            nat.check(L.lpfnt_plan_set_option(plan, b"dx_descending", 1))
        return plan

    def close(self):
        if getattr(self, "_plan", None):
            nat.lib().lpfnt_plan_destroy(self._plan)
            self._plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------ properties
    @property
    def spatial_dimension(self) -> int:
        return self._m

    @property
    def polynomial_degree(self) -> int:
        return self._n

    @property
    def lp_degree(self) -> float:
        return self._p

    @property
    def tube(self) -> np.ndarray:
        return self._T

    @property
    def multi_index_set(self) -> np.ndarray:
        return self._A

    @property
    def nodes(self) -> np.ndarray:
        return self._x

    @property
    def grid(self) -> np.ndarray:
        return self._grid

    @property
    def colex_order(self) -> np.ndarray:
        return self._colex_order

    @property
    def leja_order(self) -> np.ndarray:
        return self._leja_order

    @property
    def basis(self) -> str:
        return self._basis

    @property
    def device(self) -> int:
        return self._device

    @property
    def launch_count(self) -> int:
        """Kernel launches enqueued by this Transform's plan so far."""
        return int(nat.lib().lpfnt_plan_launch_count(self._plan))

    @property
    def device_bytes(self) -> int:
        return int(nat.lib().lpfnt_plan_device_bytes(self._plan))

    def set_option(self, name: str, value: int):
        nat.check(nat.lib().lpfnt_plan_set_option(self._plan, name.encode(), int(value)))

    def trace_read(self, max_records: int = 1 << 16):
        """Drain the per-launch CUDA-event trace (enable with set_option("trace", 1)).
        Returns a list of (label, ms); label = 100*kind + coordinate, see include/lpfnt.h."""
        labels = np.zeros(max_records, dtype=np.int32)
        ms = np.zeros(max_records, dtype=np.float32)
        cnt = C.c_int(0)
        nat.check(nat.lib().lpfnt_plan_trace_read(self._plan, max_records, labels.ctypes.data, ms.ctypes.data, C.byref(cnt)))
        k = min(cnt.value, max_records)
        return list(zip(labels[:k].tolist(), ms[:k].tolist()))

    # ------------------------------------------------------------- marshalling
    def _vec_in(self, v, what: str):
        """-> (array-or-tensor (B,N) contiguous f64, on_device, squeeze)"""
        if _is_cuda_tensor(v):
            if v.device.index != self._device:
                raise ValueError(f"{what}: tensor lives on cuda:{v.device.index}, the plan on cuda:{self._device}")
            # (a launch-bound caller -- dx in a loop, 7 us of kernel -- pays for every tensor op here: convert only when needed)
            t = v if (v.dtype is torch.float64 and v.is_contiguous()) else v.to(torch.float64).contiguous()
            squeeze = t.dim() == 1
            t = t.unsqueeze(0) if squeeze else t
            if t.dim() != 2 or t.shape[1] != self._N_0:
                raise ValueError(f"{what}: expected shape ({self._N_0},) or (B, {self._N_0}), got {tuple(v.shape)}")
            return t, True, squeeze
        a = np.ascontiguousarray(np.asarray(v), dtype=np.float64)
        squeeze = a.ndim == 1
        a = a.reshape(1, -1) if squeeze else a
        if a.ndim != 2 or a.shape[1] != self._N_0:
            raise ValueError(f"{what}: expected shape ({self._N_0},) or (B, {self._N_0}), got {np.shape(v)}")
        return a, False, squeeze

    @staticmethod
    def _ptr(x):
        return x.data_ptr() if _is_cuda_tensor(x) else x.ctypes.data

    def _stream(self, on_device: bool):
        if not on_device:
            return None
        return C.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def _out_like(self, x, on_device: bool, out=None):
        if out is not None:
            ok = (_is_cuda_tensor(out) == on_device) and tuple(out.shape)[-1] == self._N_0 and (
                out.dtype == (torch.float64 if on_device else np.float64))
            if on_device:
                ok = ok and out.is_contiguous() and out.numel() == x.numel()
            else:
                ok = ok and out.flags["C_CONTIGUOUS"] and out.size == x.size
            if not ok:
                raise ValueError("out= must be a contiguous float64 array of the input's shape on the same side (host/device)")
            return out if out.shape == x.shape else out.reshape(x.shape)
        if on_device:
            return torch.empty_like(x)
        return np.empty_like(x)

    def _finish(self, out, squeeze: bool, user_out=None):
        if user_out is not None and (user_out.ndim == 1) == squeeze:
            return user_out  # the caller's own array (same memory as `out`)
        return out[0] if squeeze else out

    # ------------------------------------------------------------------ ops
    def warmup(self) -> None:
        """Run every op once on zeros (loads the kernels; mirrors transform.py:415-428)."""
        # on DEVICE arrays: a host array would leave the two staging buffers of the host-pointer path (2 x 8 N bytes) in the
        # plan for good -- 112 MB at d = 6, n = 20 that a device-resident caller never uses
        if torch is None:
            z = np.zeros(self._N_0, dtype=np.float64)
            self.fnt(z); self.ifnt(z); self.dx(z, 0); self.dxT(z, 0)
            self.eval(z, np.zeros((1, self._m), dtype=np.float64))
            return
        dev = torch.device("cuda", self._device)
        z = torch.zeros(self._N_0, dtype=torch.float64, device=dev)
        self.fnt(z)
        self.ifnt(z)
        self.dx(z, 0)
        self.dxT(z, 0)
        self.eval(z, torch.zeros((1, self._m), dtype=torch.float64, device=dev))
        torch.cuda.synchronize(dev)

    def fnt(self, function_values, out=None):
        """Fast Newton Transform: values on ``grid`` -> Newton coefficients (transform.py:430-551).

        Bit-identical to the reference (same operation order, unfused multiply/add)."""
        x, dev, sq = self._vec_in(function_values, "fnt")
        user_out = out
        out = self._out_like(x, dev, out)
        nat.check(nat.lib().lpfnt_fnt(self._plan, self._ptr(x), self._ptr(out), x.shape[0], int(dev), self._stream(dev)))
        return self._finish(out, sq, user_out)

    def ifnt(self, coefficients, out=None):
        """Inverse transform: coefficients -> values on ``grid`` (transform.py:553-632)."""
        x, dev, sq = self._vec_in(coefficients, "ifnt")
        user_out = out
        out = self._out_like(x, dev, out)
        arith = nat.ARITH_EXACT if self._exact_ifnt else nat.ARITH_FMA
        nat.check(nat.lib().lpfnt_ifnt(self._plan, self._ptr(x), self._ptr(out), x.shape[0], arith, int(dev), self._stream(dev)))
        return self._finish(out, sq, user_out)

    def _check_ik(self, i, k):
        i, k = int(i), int(k)
        if i < 0 or i >= self._m:
            raise ValueError(f"Invalid value for i. Please choose i in between 0 and {self._m - 1}.")
        if k not in (1, 2, 3):
            raise ValueError("Invalid value for k. Please choose 1, 2 or 3.")
        return i, k

    def dx(self, coefficients, i: int, k: Literal[1, 2, 3] = 1, out=None):
        """k-th partial derivative along coordinate i, in coefficient space (transform.py:634-718)."""
        i, k = self._check_ik(i, k)
        x, dev, sq = self._vec_in(coefficients, "dx")
        user_out = out
        out = self._out_like(x, dev, out)
        nat.check(nat.lib().lpfnt_dx(self._plan, self._ptr(x), self._ptr(out), i, k, x.shape[0], int(dev), self._stream(dev)))
        return self._finish(out, sq, user_out)

    def dxT(self, coefficients, i: int, k: Literal[1, 2, 3] = 1, out=None):
        """Adjoint of :meth:`dx` (transform.py:720-804)."""
        i, k = self._check_ik(i, k)
        x, dev, sq = self._vec_in(coefficients, "dxT")
        user_out = out
        out = self._out_like(x, dev, out)
        nat.check(nat.lib().lpfnt_dxT(self._plan, self._ptr(x), self._ptr(out), i, k, x.shape[0], int(dev), self._stream(dev)))
        return self._finish(out, sq, user_out)

    def eval(self, coefficients, points):
        """Evaluate the expansion at arbitrary points, shape (P, m) -> (P,) (transform.py:806-848)."""
        dev = _is_cuda_tensor(coefficients)
        if dev != _is_cuda_tensor(points):
            if dev:
                points = torch.as_tensor(np.asarray(points, dtype=np.float64), device=coefficients.device)
            else:
                coefficients = torch.as_tensor(np.asarray(coefficients, dtype=np.float64), device=points.device)
                dev = True
        if dev:
            c = coefficients.to(torch.float64).contiguous().reshape(-1)
            p = points.to(torch.float64).contiguous()
            p = p.reshape(-1, self._m) if p.dim() != 2 else p
            out = torch.empty(p.shape[0], dtype=torch.float64, device=c.device)
        else:
            c = np.ascontiguousarray(np.asarray(coefficients), dtype=np.float64).reshape(-1)
            p = np.ascontiguousarray(np.asarray(points), dtype=np.float64)
            p = p.reshape(-1, self._m) if p.ndim != 2 else p
            out = np.empty(p.shape[0], dtype=np.float64)
        if c.shape[0] != self._N_0:
            raise ValueError(f"eval: expected {self._N_0} coefficients, got {c.shape[0]}")
        if p.shape[1] != self._m:
            raise ValueError(f"eval: points must have shape (P, {self._m})")
        nat.check(nat.lib().lpfnt_eval(self._plan, self._ptr(c), self._ptr(p), p.shape[0], self._ptr(out), int(dev), self._stream(dev)))
        return out

    def embed(self, t: "Transform") -> np.ndarray:
        """Index map of this (coarser) space into ``t`` (transform.py:850-899)."""
        if t.spatial_dimension != self.spatial_dimension:
            raise ValueError("Spatial dimensions do not match.")
        if not np.allclose(t.nodes[: self.polynomial_degree + 1], self.nodes):
            raise ValueError("Nodes mismatch: The nodes of `self` must be the starting nodes of `t`.")
        if not (t.lp_degree >= self.lp_degree and t.polynomial_degree >= self.polynomial_degree):
            raise ValueError("The index set of the transform `t` must already contain the index set of `self` for embedding.")
        return lpset.ordinal_embedding(self._A, t.multi_index_set)

    def _transfer(self, coefficients, t: "Transform", restrict: bool):
        # the index map only depends on the two spaces: computed once per partner (keyed by its defining parameters)
        cache = self.__dict__.setdefault("_phi_cache", {})
        key = (t._m, t._n, t._p, len(t), t._x.tobytes())
        phi = cache.get(key)
        if phi is None:
            phi = cache[key] = np.ascontiguousarray(self.embed(t), dtype=np.int64)
        src = t if restrict else self
        x, dev, sq = src._vec_in(coefficients, "restrict" if restrict else "prolong")
        n_out = self._N_0 if restrict else len(t)
        out = torch.empty((x.shape[0], n_out), dtype=torch.float64, device=x.device) if dev else np.empty((x.shape[0], n_out))
        nat.check(nat.lib().lpfnt_embed_apply(self._device, phi.ctypes.data, self._N_0, len(t), self._ptr(x), self._ptr(out), x.shape[0],
                                              int(restrict), int(dev), self._stream(dev)))
        return out[0] if sq else out

    def prolong(self, coefficients, t: "Transform"):
        """Coefficients of this (coarser) space as coefficients of ``t`` (zero elsewhere): ``fine[self.embed(t)] = coarse``,
        the adaptivity workflow of the reference's README, on the device.  Extension (the reference stops at ``embed``)."""
        return self._transfer(coefficients, t, False)

    def restrict(self, coefficients_fine, t: "Transform"):
        """Inverse of :meth:`prolong`: the entries of a coefficient vector of ``t`` that belong to this space."""
        return self._transfer(coefficients_fine, t, True)

    def __len__(self) -> int:
        return self._N_0

    def __eq__(self, other) -> bool:
        return (
            isinstance(other, Transform)
            and other._m == self._m
            and other._n == self._n
            and other._p == self._p
            and np.array_equal(other._x, self._x)
        )

    __hash__ = None

    def __repr__(self) -> str:
        bar = f"{'-'*20}-+-{'-'*20}\n"
        return (
            bar + f"{' '*19}Report{' '*18}\n" + bar
            + f"{'Spatial Dimension':<20} | {self._m}\n"
            + f"{'Polynomial Degree':<20} | {self._n:_}\n"
            + f"{'lp Degree':<20} | {self._p}\n"
            + f"{'Condition V':<20} | {self._cond_Vx:.2e}\n"
            + f"{'Amount of Coeffs':<20} | {len(self):_}\n"
            + f"{'Construction':<20} | {self._construction_ms:_.2f} ms\n"
            + f"{'Plan upload/warm-up':<20} | {self._precompilation_ms:_.2f} ms\n"
            + f"{'Device':<20} | cuda:{self._device}\n"
            + bar
        )

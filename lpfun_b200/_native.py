"""ctypes binding of liblpfnt.so (the C ABI in include/lpfnt.h).

There is NO fallback: if the library is missing or an entry point fails, an exception is
raised.  The compute path of this package is the CUDA library and nothing else.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblpfnt.so")

LPFNT_LOWER, LPFNT_UPPER = 0, 1
ARITH_EXACT, ARITH_FMA = 0, 1

_lib = None


class LpfntError(RuntimeError):
    pass


def _declare(L):
    i64, i32, dbl, vp = C.c_int64, C.c_int, C.c_double, C.c_void_p
    L.lpfnt_version.restype = C.c_char_p
    L.lpfnt_last_error.restype = C.c_char_p
    L.lpfnt_device_count.argtypes = [C.POINTER(i32)]
    L.lpfnt_lp_set_count.restype = i64
    L.lpfnt_lp_set_count.argtypes = [i32, i32, dbl]
    L.lpfnt_lp_set_fill.argtypes = [i32, i32, dbl, vp, i64]
    L.lpfnt_lp_tube.restype = i64
    L.lpfnt_lp_tube.argtypes = [vp, i64, i32, vp]
    L.lpfnt_entropy.argtypes = [vp, i64, i32, vp]
    L.lpfnt_ordinal_embedding.argtypes = [vp, i64, vp, i64, i32, vp]
    L.lpfnt_index_create.argtypes = [vp, i64, i32, i32, C.POINTER(vp)]
    L.lpfnt_index_destroy.argtypes = [vp]
    L.lpfnt_index_size.restype = i64
    L.lpfnt_index_size.argtypes = [vp, i32, i32]
    L.lpfnt_index_copy.argtypes = [vp, i32, i32, vp]
    L.lpfnt_index_whole_build.argtypes = [vp, i32, i32, i32]
    L.lpfnt_index_whole_meta.argtypes = [vp, vp]
    L.lpfnt_index_wsplit_build.argtypes = [vp, i32, i32]
    L.lpfnt_index_wsplit_meta.argtypes = [vp, vp]
    L.lpfnt_host_alloc.argtypes = [C.POINTER(vp), i64]
    L.lpfnt_host_free.argtypes = [vp]
    L.lpfnt_plan_create.argtypes = [i32, i32, i64, vp, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(vp), i32, C.POINTER(vp)]
    L.lpfnt_plan_set_upper_factors.argtypes = [vp, vp, vp, i32]
    L.lpfnt_sweep.argtypes = [vp, vp, i32, i32, i32, C.c_uint32, vp, vp, i64, i32, vp]
    L.lpfnt_embed_apply.argtypes = [i32, vp, i64, i64, vp, vp, i64, i32, i32, vp]
    L.lpfnt_plan_destroy.argtypes = [vp]
    L.lpfnt_plan_set_push.argtypes = [vp, vp, i64, i32, C.POINTER(vp), vp]
    L.lpfnt_push_rows.argtypes = [i32, vp, i64, i32, C.POINTER(vp), vp, vp]
    L.lpfnt_push_rows_ex.argtypes = [i32, vp, i64, i32, C.POINTER(vp), vp, i32, vp]
    L.lpfnt_push_rows_ce.argtypes = [i32, vp, i64, i32, C.POINTER(vp), vp]
    L.lpfnt_plan_device_bytes.restype = i64
    L.lpfnt_plan_device_bytes.argtypes = [vp]
    L.lpfnt_plan_launch_count.restype = i64
    L.lpfnt_plan_launch_count.argtypes = [vp]
    L.lpfnt_plan_set_option.argtypes = [vp, C.c_char_p, i64]
    L.lpfnt_plan_trace_read.argtypes = [vp, i32, vp, vp, C.POINTER(i32)]
    L.lpfnt_plan_timeline_read.argtypes = [vp, vp, i32]
    L.lpfnt_itransform.argtypes = [vp, vp, i32, i32, vp, vp, i64, i32, i32, i32, vp]
    L.lpfnt_transform.argtypes = [vp, vp, i32, vp, vp, i64, i32, i32, i32, vp]
    L.lpfnt_dtransform.argtypes = [vp, vp, i32, i32, i32, vp, vp, i64, i32, vp]
    L.lpfnt_newton2point.argtypes = [vp, vp, vp, i64, vp, i32, vp]
    L.lpfnt_fnt.argtypes = [vp, vp, vp, i64, i32, vp]
    L.lpfnt_ifnt.argtypes = [vp, vp, vp, i64, i32, i32, vp]
    L.lpfnt_dx.argtypes = [vp, vp, vp, i32, i32, i64, i32, vp]
    L.lpfnt_dxT.argtypes = [vp, vp, vp, i32, i32, i64, i32, vp]
    L.lpfnt_eval.argtypes = [vp, vp, vp, i64, vp, i32, vp]
    L.lpfnt_measure_fp64.argtypes = [i32, i32, C.POINTER(dbl)]
    return L


def lib():
    """Load liblpfnt.so; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m lpfun_b200._build` "
                "(nvcc, sm_100a). lpfun_b200 has no CPU fallback."
            )
        _lib = _declare(C.CDLL(LIB_PATH))
    return _lib


def last_error() -> str:
    return lib().lpfnt_last_error().decode("utf-8", "replace")


def check(rc: int):
    if rc == 0:
        return
    msg = last_error()
    if rc == -1:
        raise ValueError(msg)
    if rc == -4:
        raise NotImplementedError(msg)
    raise LpfntError(f"liblpfnt error {rc}: {msg}")


def ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().lpfnt_device_count(C.byref(n))
    return int(n.value) if rc == 0 else 0


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """NumPy array over page-locked host memory (cudaHostAlloc), for the host-pointer path."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(lib().lpfnt_host_alloc(C.byref(p), nbytes))
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.__array_interface__["data"][0]] = p.value
    return arr


_PINNED: dict[int, int] = {}


def pinned_free(arr: np.ndarray):
    addr = arr.__array_interface__["data"][0]
    p = _PINNED.pop(addr, None)
    if p is not None:
        check(lib().lpfnt_host_free(p))


(TAB_LINE_OFF, TAB_FIB_PTR, TAB_FIB_ENT, TAB_THR_START, TAB_THR_FIB, TAB_LINE_ALPHA, TAB_LEVEL_SIZE,
 TAB_GROUPS, TAB_LINE_ORDER, TAB_GROUP_T, TAB_WHOLE_ITEMS, TAB_WHOLE_SEO, TAB_LINE_LAST, TAB_WSPLIT_ITEMS, TAB_WSPLIT_SEO) = range(15)
_TAB_DTYPE = {TAB_LINE_ALPHA: np.uint8, TAB_LEVEL_SIZE: np.int64, TAB_GROUP_T: np.uint8, TAB_WHOLE_SEO: np.uint16, TAB_LINE_LAST: np.uint8,
              TAB_WHOLE_ITEMS: np.uint32, TAB_WSPLIT_ITEMS: np.uint32, TAB_WSPLIT_SEO: np.uint16}


class IndexTables:
    """Host copy of the derived device tables of a multi-index set (no GPU needed)."""

    def __init__(self, A: np.ndarray, n: int):
        A = np.ascontiguousarray(A, dtype=np.int64)
        self.m = A.shape[1]
        self._h = C.c_void_p()
        check(lib().lpfnt_index_create(A.ctypes.data, A.shape[0], self.m, int(n), C.byref(self._h)))

    def table(self, what: int, h: int = 0) -> np.ndarray:
        cnt = lib().lpfnt_index_size(self._h, what, h)
        if cnt < 0:
            check(int(cnt))
        out = np.empty(cnt, dtype=_TAB_DTYPE.get(what, np.int32))
        if cnt:
            check(lib().lpfnt_index_copy(self._h, what, h, out.ctypes.data))
        return out.reshape(-1, 2) if what in (TAB_FIB_ENT, TAB_GROUPS) else out

    def whole_build(self, threads: int = 768, order: int = -1, fibres: int = 1):
        check(lib().lpfnt_index_whole_build(self._h, int(threads), int(order), int(fibres)))

    def wsplit_build(self, C: int = 0, order: int = -1):
        check(lib().lpfnt_index_wsplit_build(self._h, int(C), int(order)))

    def wsplit_meta(self) -> dict:
        meta = np.zeros(83, dtype=np.int32)
        lib().lpfnt_index_wsplit_meta(self._h, meta.ctypes.data)
        d = {"ok": int(meta[0]), "C": int(meta[1]), "e0": meta[[2, 4]].tolist(), "ne": meta[[3, 5]].tolist()}
        lists = meta[6:18].reshape(2, 3, 2)
        d["item_off"] = lists[:, :, 0].tolist()
        d["n_items"] = lists[:, :, 1].tolist()
        d["ncolq"] = int(meta[18])
        d["cb"] = meta[19:51].tolist()
        d["rq"] = meta[51:83].tolist()
        return d

    def whole_meta(self) -> dict:
        meta = np.zeros(10, dtype=np.int32)
        lib().lpfnt_index_whole_meta(self._h, meta.ctypes.data)
        d = dict(zip(["ok", "threads", "max_t", "max_items"], meta[:4].tolist()))
        lists = meta[4:10].reshape(3, 2)
        d["item_off"] = lists[:, 0].tolist()
        d["n_items"] = lists[:, 1].tolist()
        return d

    def close(self):
        if self._h:
            lib().lpfnt_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

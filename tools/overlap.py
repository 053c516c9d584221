"""One GPU: does the streaming copy kernel (lpfnt_push_rows) run BESIDE the persistent sweep kernels?  Times the sweeps of
one chunk, the copy of one chunk into another local buffer, and both on two streams."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpfun_b200  # noqa: E402
from lpfun_b200 import _native as nat  # noqa: E402

B = int(os.environ.get("B", 888))
t = lpfun_b200.Transform(3, 30, report=False)
N = len(t)
dev = torch.device("cuda:0")
vals = torch.rand((B, N), dtype=torch.float64, device=dev)
coef = torch.empty_like(vals); rec = torch.empty_like(vals)
CB = int(os.environ.get("CB", 4 * B))  # rows copied: several chunks' worth, so that the copy lasts about as long as the sweeps
src = torch.rand((CB, N), dtype=torch.float64, device=dev)
dst = torch.empty_like(src)
side = torch.cuda.Stream(dev)
engine = os.environ.get("ENGINE", "sm")


def sweeps():
    t.fnt(vals, out=coef); t.ifnt(coef, out=rec)


def copy(stream):
    one = (C.c_void_p * 1)(dst.data_ptr())
    if engine == "ce":
        nat.check(nat.lib().lpfnt_push_rows_ce(0, C.c_void_p(src.data_ptr()), src.numel() * 8, 1, one, C.c_void_p(stream.cuda_stream)))
    else:
        nat.check(nat.lib().lpfnt_push_rows_ex(0, C.c_void_p(src.data_ptr()), src.numel() * 8, 1, one, C.c_void_p(0), int(os.environ.get('CTAS', 1)), C.c_void_p(stream.cuda_stream)))


def timed(f, K=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


def both(copy_first):
    cur = torch.cuda.current_stream(dev)
    side.wait_stream(cur)
    if copy_first:
        copy(side); sweeps()
    else:
        sweeps(); copy(side)
    cur.wait_stream(side)


a = timed(sweeps)
b = timed(lambda: copy(torch.cuda.current_stream(dev)))
c = timed(lambda: both(True))
d = timed(lambda: both(False))
assert torch.equal(src, dst)
print(f"engine {engine} B={B} CB={CB} ctas/SM={os.environ.get('CTAS','1')} carveout={os.environ.get('LPFNT_DEBUG_PUSH_CARVEOUT','-')}: "
      f"sweeps {a:.3f} ms, copy {b:.3f} ms ({src.numel()*8/b/1e6:.0f} GB/s), both(copy launched first) {c:.3f} ms, both(sweeps first) {d:.3f} ms")

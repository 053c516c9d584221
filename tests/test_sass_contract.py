"""The instruction contract behind the bit-exactness claims, checked on the built library (no GPU needed):
the reference's fnt is only reproducible with separately rounded multiplies and adds, so no
EXACT instantiation of a mat-vec sweep kernel may contain a fused multiply-add -- a compiler upgrade that contracts
__dmul_rn/__dadd_rn, or a leaked -fmad, fails here, not in a parity test on some input."""
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_hist  # noqa: E402

KERNEL = re.compile(r"lpfnt::(k_whole|k_slab|k_sweep_lines|k_sweep_groups)<(\d+), (\d+), (true|false)")
MATVEC_MODES = {0, 1, 5}  # lower / upper / descending upper mat-vec (the solves divide: __ddiv_rn expands to an FMA sequence)


@pytest.fixture(scope="module")
def hist(built_lib):
    if sass_hist.tool("cuobjdump") is None:
        pytest.skip("cuobjdump not installed")
    from lpfun_b200 import _build

    h = sass_hist.histogram(_build.LIB)
    names = sass_hist.demangle(list(h))
    return {names[k]: v for k, v in h.items()}


def count(c, op):
    return sum(v for k, v in c.items() if k == op or k.startswith(op + "."))


def test_exact_kernels_never_fuse(hist):
    seen = 0
    for name, c in hist.items():
        m = KERNEL.search(name)
        if not m or int(m.group(3)) not in MATVEC_MODES:
            continue
        fma = m.group(4) == "true"
        if fma:
            assert count(c, "DFMA") > 0, name
        else:
            seen += 1
            assert count(c, "DFMA") == 0, f"{name}: {count(c, 'DFMA')} DFMA in an exact instantiation"
            assert count(c, "DMUL") > 0 and count(c, "DADD") > 0, name
    assert seen >= 20, f"only {seen} exact mat-vec instantiations found"


def test_whole_function_kernel_uses_the_bulk_copy_engine(hist):
    whole = [c for n, c in hist.items() if "lpfnt::k_whole<" in n]
    assert whole
    for c in whole:
        assert count(c, "UBLKCP") > 0 and count(c, "SYNCS") > 0  # cp.async.bulk + mbarrier


def test_no_tensor_core_fp64_on_the_path(hist):
    # DMMA runs at 80 % of the DFMA rate on B200 and is FMA-based; nothing on this path uses it
    assert all(count(c, "DMMA") == 0 for c in hist.values())

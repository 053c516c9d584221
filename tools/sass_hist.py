#!/usr/bin/env python
"""Opcode histogram of the kernels in liblpfnt.so (cuobjdump -sass): the instruction contract behind the bit-exactness
claims -- the exact instantiations (template argument FMA = false) of the mat-vec sweeps contain DMUL + DADD and not one
DFMA; the whole-function kernel fetches with the bulk-copy engine (UBLKCP).
    python tools/sass_hist.py [liblpfnt.so] > profiles/sass/<name>.txt
Used by tests/test_sass_contract.py."""
import collections
import os
import re
import shutil
import subprocess
import sys

OPS = ("DFMA", "DMUL", "DADD", "DMMA", "UBLKCP", "LDCU", "LDS", "STS", "LDG", "STG", "LDGSTS", "BAR", "SYNCS", "MUFU")
LINE = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")


def tool(name):
    for cand in (shutil.which(name), f"/usr/local/cuda/bin/{name}"):
        if cand and os.path.exists(cand):
            return cand
    return None


def histogram(lib):
    """{mangled function name: Counter(opcode -> static count)} for every kernel in `lib`."""
    cuobjdump = tool("cuobjdump")
    if cuobjdump is None:
        raise FileNotFoundError("cuobjdump")
    proc = subprocess.Popen([cuobjdump, "-sass", lib], stdout=subprocess.PIPE, text=True)
    out, cur = {}, None
    for line in proc.stdout:
        if "Function :" in line:
            cur = out.setdefault(line.split("Function :")[1].strip(), collections.Counter())
            continue
        if cur is None:
            continue
        m = LINE.match(line)
        if m:
            op = m.group(1)
            cur[op] += 1
    proc.wait()
    return out


def demangle(names):
    filt = tool("c++filt")
    if filt is None:
        return {n: n for n in names}
    res = subprocess.run([filt], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, res))


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "..", "lpfun_b200", "liblpfnt.so")
    h = histogram(lib)
    names = demangle(list(h))
    print(f"# {len(h)} kernels in {os.path.basename(lib)}; static opcode counts (cuobjdump -sass)")
    print("# " + "  ".join(f"{o:>6s}" for o in OPS) + "  kernel")
    for mangled in sorted(h, key=lambda k: names[k]):
        c = h[mangled]
        short = re.sub(r"\(.*", "", names[mangled]).replace("lpfnt::", "").replace("void ", "")
        print("  " + "  ".join(f"{sum(v for k, v in c.items() if k == o or k.startswith(o + '.')):6d}" for o in OPS) + "  " + short)


if __name__ == "__main__":
    main()

"""In-tree build of liblpfnt.so (nvcc, sm_100a only).

``python -m lpfun_b200._build`` or ``lpfun_b200._build.build()``.  Objects are compiled in
parallel into ``lpfun_b200/csrc/_obj`` and linked to ``lpfun_b200/liblpfnt.so`` next to the
package, so the library travels with the source tree (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "liblpfnt.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
GROUP_LANES = os.environ.get("LPFNT_GROUP_LANES")
COMMON = ([f"-DLPFNT_GROUP_LANES={GROUP_LANES}"] if GROUP_LANES else []) + (["-DLPFNT_TIMELINE_SPLIT"] if os.environ.get("LPFNT_TIMELINE_SPLIT") else []) + ["-O3", "-std=c++17", "-lineinfo", "-Xfatbin=-compress-all", "-Xcompiler", "-fPIC"]

# (source, object name, extra defines)
UNITS = [
    ("index_set.cpp", "index_set.o", []),
    ("api.cu", "api.o", []),
    ("eval_inst.cu", "eval_inst.o", []),
] + [("sweep_inst.cu", f"sweep_T{t}.o", [f"-DLPFNT_MAXT={t}"]) for t in (8, 16, 24, 32)]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def _host_compiler() -> list[str]:
    # the image exports CC/CXX=/opt/gcc/bin/*, wrappers that break some flags: prefer the system g++
    for cand in ("/usr/bin/g++",):
        if os.path.exists(cand):
            return ["-ccbin", cand]
    return []


def _deps_mtime() -> float:
    latest = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in os.listdir(root):
            if name.endswith((".h", ".cuh", ".cu", ".cpp")):
                latest = max(latest, os.path.getmtime(os.path.join(root, name)))
    return latest


def build(force: bool = False, verbose: bool = False) -> str:
    units = list(UNITS)
    newest = _deps_mtime()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    ccbin = _host_compiler()

    def compile_one(unit):
        src, obj, defs = unit
        out = os.path.join(OBJ, obj)
        if not force and os.path.exists(out) and os.path.getmtime(out) >= newest:
            return out, ""
        cmd = [nvcc, *ccbin, *ARCH, *COMMON, *defs, "-c", os.path.join(CSRC, src), "-o", out]
        if src.endswith(".cpp"):
            cmd.insert(1, "-x")
            cmd.insert(2, "cu")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        return out, r.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, len(units))) as ex:
        results = list(ex.map(compile_one, units))
    objs = [r[0] for r in results]
    if verbose:
        for _, log in results:
            if log.strip():
                print(log, file=sys.stderr)
    link = [nvcc, *ccbin, *ARCH, "-shared", "-Xlinker", "-s", "-o", LIB, *objs]  # -s: no host symbol table (the exported C-ABI stays in .dynsym)
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{' '.join(link)}\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

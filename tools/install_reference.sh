#!/bin/bash
# Puts the UNMODIFIED reference package where bench.py --impl reference looks first (baseline/_ref/src), so that it
# travels to the GPU box with the snapshot (baseline/_ref is git-ignored, not gpurun-ignored).
# `pip install --no-index --no-build-isolation --target baseline/_ref /root/reference` fails in this image
# (the package's build backend, poetry-core, is not in the offline wheelhouse); the package is pure Python
# (src/lpfun/*.py), so the install step is a plain copy of its source tree.
set -e
SRC=${1:-/root/reference/src}
DST="$(dirname "$0")/../baseline/_ref/src"
mkdir -p "$DST"
rm -rf "$DST/lpfun"
cp -r "$SRC/lpfun" "$DST/lpfun"
find "$DST" -name __pycache__ -type d -exec rm -rf {} + 2>/dev/null || true
echo "installed $(ls "$DST/lpfun" | wc -l) entries into $DST/lpfun"

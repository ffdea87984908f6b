#!/bin/sh
# Stage the UNMODIFIED reference (its two Python packages and the three model YAMLs) under the git-ignored
# baseline/_ref/ so that the reference's own model and wrappers can run on the GPU box next to the B200 backend
# (tools/model_bench.py, the wrapper / model tests in tests/test_recurrent_and_wrappers_gpu.py, bench.py's
# model_train block).  gpurun ships baseline/_ref with the snapshot; git never sees it (.gitignore).  Everything that
# reads it skips or reports "unavailable" when it is absent; the product never imports it.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
DST="$ROOT/baseline/_ref"
[ -d "$SRC/mlstm_kernels" ] || { echo "no reference at $SRC" >&2; exit 1; }
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/mlstm_kernels" "$SRC/ultralytics" "$DST/"
cp "$SRC"/640-base*.yaml "$DST/"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
echo "staged $(du -sh "$DST" | cut -f1) into $DST"

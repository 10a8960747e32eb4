#!/usr/bin/env bash
# Install the UNMODIFIED reference into the git-ignored oracle/_ref/ so that it can travel to the GPU
# box with the working tree (like the built .so files): the reference is pure Python, `pip install`
# needs the build backend `hatchling`, which the offline wheelhouse does not have, so the package
# directory is copied as is.  Next to it go the import shims of tools/refshim (mpi4py, cerberus,
# colorlog, colorama, h5py, matplotlib, stl: none carries arithmetic).
#
#   tools/make_ref.sh [reference root, default /root/reference]
#
# Result:  oracle/_ref/pylbm/          the reference package, byte-identical (checked with diff -r)
#          oracle/_ref/shims/          copy of tools/refshim
#          oracle/_ref/MANIFEST.txt    sha256 of every copied reference file
# Users: tests (plugin parity), bench.py's reference arm / cpu_baseline (generator='cython' timed on
# the host) and the plugin path of bench.py (the reference's own front-end drives the CUDA kernels).
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
REF="${1:-${PYLBM_REFERENCE:-/root/reference}}"
OUT="$ROOT/oracle/_ref"
if [ ! -d "$REF/pylbm" ]; then
    echo "make_ref: no reference at $REF (keeping what is in $OUT)" >&2
    exit 0
fi
rm -rf "$OUT.tmp"
mkdir -p "$OUT.tmp"
cp -r "$REF/pylbm" "$OUT.tmp/pylbm"
cp -r "$ROOT/tools/refshim" "$OUT.tmp/shims"
find "$OUT.tmp" -name '__pycache__' -type d -prune -exec rm -rf {} +
diff -r -x '__pycache__' "$REF/pylbm" "$OUT.tmp/pylbm" > /dev/null
(cd "$OUT.tmp" && find pylbm -type f | sort | xargs sha256sum) > "$OUT.tmp/MANIFEST.txt"
rm -rf "$OUT"
mv "$OUT.tmp" "$OUT"
echo "make_ref: $(find "$OUT/pylbm" -name '*.py' | wc -l) python files -> $OUT"

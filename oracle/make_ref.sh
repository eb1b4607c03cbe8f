#!/bin/sh
# Stage the UNMODIFIED reference package (Orange-OpenSource/Cool-Chic, pure Python) into oracle/_ref/ so that
# bench.py --impl reference can time the reference's own decode_video on the GPU box's host cores.
# oracle/_ref/ is git-ignored (never part of the history) but NOT gpurun-ignored (it travels with the snapshot).
# The two third-party imports the reference needs and this image lacks (constriction 0.4.2, fvcore) come from
# oracle/refshim/ (constriction's range coder forwards to oracle/_build/libccoracle.so).
# TEST / BENCHMARK INFRASTRUCTURE ONLY: nothing in the product imports it.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
if [ ! -d "$SRC/coolchic" ]; then
    echo "make_ref: $SRC/coolchic not found (nothing staged)" >&2
    exit 0
fi
rm -rf "$HERE/_ref/coolchic"
mkdir -p "$HERE/_ref"
cp -r "$SRC/coolchic" "$HERE/_ref/coolchic"
find "$HERE/_ref" -name __pycache__ -type d -exec rm -rf {} + 2>/dev/null || true
( cd "$SRC" && find coolchic -type f -name '*.py' | sort | xargs sha256sum ) > "$HERE/_ref/SHA256SUMS"
echo "staged $(find "$HERE/_ref/coolchic" -name '*.py' | wc -l) reference files into oracle/_ref/"

#!/usr/bin/env bash
# Stage the UNMODIFIED reference hot path next to the oracle so that it can travel to the GPU box
# (/root/reference does not exist there).  Output only into oracle/_ref/ (git-ignored, NOT gpurun-ignored):
#   oracle/_ref/ficp.py                         <- /root/reference/ficp.py            (the CPU arm of bench.py)
#   oracle/_ref/tests/test_ficp.py              <- /root/reference/tests/test_ficp.py (run verbatim against this repo's ficp)
#   oracle/_ref/tests/test_rigid_2d_operations.py
#   oracle/_ref/trees.py                        <- imported by test_rigid_2d_operations.py (Plot rigid edits)
#   oracle/_ref/SHA256SUMS                      <- checked by tests/test_reference_vendored.py
# Nothing is edited; the files are byte copies (cmp below).  Run by __graft_entry__.build() when /root/reference exists.
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$HERE/oracle/_ref"
[ -f "$REF/ficp.py" ] || { echo "vendor_ref: $REF/ficp.py not found (GPU box: use the prebuilt oracle/_ref)"; exit 0; }
mkdir -p "$OUT/tests"
for f in ficp.py trees.py tests/test_ficp.py tests/test_rigid_2d_operations.py; do
    cp "$REF/$f" "$OUT/$f"
    cmp -s "$REF/$f" "$OUT/$f"
done
(cd "$OUT" && sha256sum ficp.py trees.py tests/test_ficp.py tests/test_rigid_2d_operations.py > SHA256SUMS)
echo "vendor_ref: staged $(wc -l < "$OUT/SHA256SUMS") files under $OUT"

#!/usr/bin/env bash
# Copy the UNMODIFIED reference (Python sources of the scoring path + its sample data) to baseline/_ref/.
# baseline/_ref/ is git-ignored (never part of the history) but NOT gpurun-ignored, so it travels to the GPU box,
# where oracle/ref_harness.py imports it for: bench.py --impl reference (CPU arm, kind "reference"), bench.py's
# gpu_reference leg (reference bf16 + flash-attn 2 on the same B200) and tests/test_reference_gpu.py.
# Runs only in the build container (/root/reference does not exist on the GPU box).
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="${1:-/root/reference}"
DST="$ROOT/baseline/_ref"
[ -d "$SRC/llava_reward" ] || { echo "no reference at $SRC" >&2; exit 1; }
mkdir -p "$DST/data"
rm -rf "$DST/llava_reward" "$DST/eval" "$DST/data/sample_test"
cp -r "$SRC/llava_reward" "$DST/llava_reward"
cp -r "$SRC/eval" "$DST/eval"
cp -r "$SRC/data/sample_test" "$DST/data/sample_test"
cp "$SRC/LICENSE" "$DST/LICENSE" 2>/dev/null || true
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
chmod -R u+w "$DST"
( cd "$SRC" && find llava_reward eval data/sample_test -type f ! -path '*/__pycache__/*' -print0 | sort -z | xargs -0 sha1sum ) > "$DST/SHA1SUMS"
echo "copied $(wc -l < "$DST/SHA1SUMS") files to $DST"

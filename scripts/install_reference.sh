#!/bin/sh
# Install the UNMODIFIED reference (annekegh/mcdiff, pure Python) under baseline/_ref so that
#   - bench.py --impl reference and bench.py's cpu_baseline time mcdiff.mc.do_mc_cycles itself (lib/mc.py:20-51),
#   - tests/test_gpu_reference_patch.py can apply INTEGRATION.md's patch to the real MCState.
# baseline/_ref is git-ignored (no reference source enters the history) but travels to the GPU box.
# /root/reference is read-only and setup.py writes build/ next to itself, hence the copy under /tmp.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
[ -d "$REF/lib" ] || { echo "no reference at $REF" >&2; exit 1; }
TMP="$(mktemp -d /tmp/mcdiff_ref.XXXXXX)"
cp -r "$REF/." "$TMP/"
rm -rf "$ROOT/baseline/_ref"
python -m pip install --quiet --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
    --target "$ROOT/baseline/_ref" "$TMP"
rm -rf "$TMP"
python - <<PY
import sys
sys.path.insert(0, "$ROOT/baseline/_ref")
import mcdiff.mc, mcdiff.MCState, mcdiff.utils
print("reference installed:", mcdiff.__file__)
PY

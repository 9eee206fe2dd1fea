#!/bin/bash
# TEST INFRASTRUCTURE ONLY.
# Builds the UNMODIFIED reference (Cython extensions + vendored hcephes) in a scratch
# directory OUTSIDE the repo so that tests/golden/make_golden.py can import it in this
# container and write golden vectors. Nothing from /root/reference is copied into the repo.
# Recipe follows SURVEY.md §8(c).
set -euo pipefail
REF=${REF:-/root/reference}
OUT=${1:-/tmp/fpt_pyref}
HERE="$(cd "$(dirname "$0")" && pwd)"
rm -rf "$OUT"; mkdir -p "$OUT"
cp -r "$REF/footprint_tools" "$REF/hcephes" "$REF/setup.py" "$REF/README.md" "$REF/MANIFEST.in" "$OUT/"
cd "$OUT"
PYTHONPATH="$HERE/pyref_stubs" python setup.py -q egg_info build_clib build_ext --inplace >build.log 2>&1 || { tail -30 build.log; exit 1; }
echo "pyref built in $OUT ; import with PYTHONPATH=$HERE/pyref_stubs:$OUT"
# Optional second argument: install the importable package (python modules + built extensions + egg-info, no build
# tree) into that directory — bench.py's reference leg uses <repo>/baseline/_ref (git-ignored; the base contract's
# location for the unmodified reference; it travels to the GPU box like the other built files).
if [ -n "${2:-}" ]; then
  mkdir -p "$2"; rm -rf "$2/footprint_tools" "$2"/footprint_tools.egg-info
  cp -r "$OUT/footprint_tools" "$2/"; cp -r "$OUT"/footprint_tools.egg-info "$2/" 2>/dev/null || true
  find "$2" -name "*.c" -delete; find "$2" -name "__pycache__" -prune -exec rm -rf {} +
  echo "installed into $2"
fi

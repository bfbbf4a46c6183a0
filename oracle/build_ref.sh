#!/usr/bin/env bash
# Builds the reference's own scoring code (simd_dot, Compare_Face_*) from the sources where they
# lie under /root/reference into oracle/_ref/ (git-ignored, travels to the GPU box).
# TEST INFRASTRUCTURE ONLY.  No reference source is written into the repository: the selected
# line ranges are piped straight into g++.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}/feature_extraction/c_version/Feature.hpp"
OUT="$HERE/_ref"
if [ ! -f "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
gen() {
  cat "$HERE/ref_prelude.hpp"
  sed -n '33,103p' "$REF"          # class Person
  cat "$HERE/ref_types.hpp"
  sed -n '273,293p' "$REF"         # simd_dot
  sed -n '345,392p' "$REF"         # Compare_Face_From_DB(FVV&, float, float*)
  sed -n '763,804p' "$REF"         # Compare_Face_Person, Compare_Face_DB
  cat "$HERE/ref_exports.cpp"
}
# serial build = deterministic semantics (the omp pragmas are ignored without -fopenmp)
gen | g++ -O2 -fPIC -shared -w -x c++ - -o "$OUT/libfeature_ref.so"
# OpenMP build = the reference as deployed (racy `sim`); used for CPU timing only
gen | g++ -O2 -fopenmp -fPIC -shared -w -x c++ - -o "$OUT/libfeature_ref_omp.so"
echo "built $OUT/libfeature_ref.so $OUT/libfeature_ref_omp.so"

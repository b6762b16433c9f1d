#!/usr/bin/env bash
# Conformance build (TEST INFRASTRUCTURE): compiles the reference's own harness
# (utils/src/*), its four result-pinning gtest suites and
# benchmarks/manual_benchmark.cu UNCHANGED, from where they lie under
# $REFERENCE, against this repository's drop-in (`cuembed::hdrs` of the
# top-level CMakeLists.txt) with the reference's vendored abseil / gtest.
# manual_benchmark_ref is the same benchmark source against the reference's own
# headers (the other arm of the same-box comparison).
# Outputs: oracle/_ref/conformance/{test_embedding_*,manual_benchmark,manual_benchmark_ref}
# (git-ignored, travels to the GPU box).  Nothing of the reference is copied.
#
#   oracle/conformance.sh [build-dir]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
REFERENCE="${REFERENCE:-/root/reference}"
BUILD="${1:-/tmp/cuembed_b200_conformance}"
OUT="$HERE/_ref/conformance"
if [ ! -d "$REFERENCE/tests" ]; then
  echo "reference tree absent: keeping prebuilt $OUT (if any)"; exit 0
fi
python -m cuembed_b200.build >/dev/null
mkdir -p "$BUILD" "$OUT"
cmake -S "$ROOT" -B "$BUILD" -G Ninja -DCMAKE_BUILD_TYPE=Release \
      -DCUEMBED_USE_PREBUILT_LIB=ON \
      -DCUEMBED_CONFORMANCE_REFERENCE_DIR="$REFERENCE" >"$BUILD/configure.log" 2>&1 \
  || { tail -40 "$BUILD/configure.log"; exit 1; }
ninja -C "$BUILD" test_embedding_forward test_embedding_transpose \
      test_embedding_backward test_embedding_against_cpu manual_benchmark manual_benchmark_ref \
      >"$BUILD/build.log" 2>&1 || { tail -60 "$BUILD/build.log"; exit 1; }
cp "$BUILD"/bin/test_embedding_* "$BUILD"/bin/manual_benchmark "$BUILD"/bin/manual_benchmark_ref "$OUT"/
echo "built $(ls "$OUT" | tr '\n' ' ')into $OUT"

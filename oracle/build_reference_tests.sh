#!/usr/bin/env bash
# oracle/build_reference_tests.sh -- compiles the reference's OWN test and benchmark sources, unmodified, against the
# cuzk_b200 host layer (TEST INFRASTRUCTURE; outputs only under oracle/_ref/bin/).
#
# It lays out a throw-away view of the reference tree under $FARM in which every file is a symlink:
#   src/common, src/poseidon/{field_arithmetic,poseidon}.{hpp,cpp}, src/merkle_tree/merkle_tree.{hpp,cpp},
#   src/*/test/*.cpp                                  -> /root/reference        (CPU implementation = oracle, tests)
#   src/poseidon/cuda/*, src/merkle_tree/merkle_tree_cuda*  -> cuzk_b200/host/src     (our replacement of the CUDA side)
# i.e. exactly the file swap INTEGRATION.md describes.  Nothing is copied; googletest is replaced by tests/cpp/gtest_shim
# (the reference's CMake downloads googletest, which cannot work offline).  Only runs where /root/reference exists.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF:-/root/reference}"
FARM="${FARM:-/tmp/cuzk_dropin_farm}"
OUT="$HERE/_ref/bin"
CXX="${CXX:-g++}"
if [ ! -d "$REF/src" ]; then echo "build_reference_tests: $REF absent; keeping prebuilt binaries if any"; exit 0; fi

rm -rf "$FARM"
mkdir -p "$FARM/src/poseidon/cuda" "$FARM/src/poseidon/test" "$FARM/src/merkle_tree/test" "$OUT"
ln -s "$REF/src/common" "$FARM/src/common"
for f in field_arithmetic.hpp field_arithmetic.cpp poseidon.hpp poseidon.cpp; do ln -s "$REF/src/poseidon/$f" "$FARM/src/poseidon/$f"; done
for f in merkle_tree.hpp merkle_tree.cpp; do ln -s "$REF/src/merkle_tree/$f" "$FARM/src/merkle_tree/$f"; done
for f in "$REF"/src/poseidon/test/*.cpp; do ln -s "$f" "$FARM/src/poseidon/test/$(basename "$f")"; done
for f in "$REF"/src/merkle_tree/test/*.cpp; do ln -s "$f" "$FARM/src/merkle_tree/test/$(basename "$f")"; done
ln -s "$REF/src/poseidon/cuda/poseidon_cuda_profiler.cpp" "$FARM/src/poseidon/cuda/poseidon_cuda_profiler.cpp"
HOST="$ROOT/cuzk_b200/host/src"
for f in "$HOST"/poseidon/cuda/*; do ln -s "$f" "$FARM/src/poseidon/cuda/$(basename "$f")"; done
for f in "$HOST"/merkle_tree/merkle_tree_cuda*; do ln -s "$f" "$FARM/src/merkle_tree/$(basename "$f")"; done

S="$FARM/src"
CPU_SRCS="$S/poseidon/field_arithmetic.cpp $S/poseidon/poseidon.cpp $S/merkle_tree/merkle_tree.cpp"
GPU_SRCS="$S/poseidon/cuda/field_arithmetic_cuda.cpp $S/poseidon/cuda/poseidon_cuda.cpp $S/poseidon/cuda/poseidon_cuda_benchmarks.cpp \
          $S/poseidon/cuda/poseidon_cuda_vs_cpu.cpp $S/merkle_tree/merkle_tree_cuda.cpp $S/merkle_tree/merkle_tree_cuda_vs_cpu.cpp"
FLAGS="-O2 -std=c++17 -I$ROOT/include -I$ROOT/tests/cpp/gtest_shim -I$S/poseidon -I$S/merkle_tree"
LINK="-L$ROOT/cuzk_b200 -lcuzk_b200 -Wl,-rpath,\$ORIGIN/../../../cuzk_b200 -lpthread"

# one object set shared by all binaries
OBJ="$FARM/obj"; mkdir -p "$OBJ"
objs=""
for src in $CPU_SRCS $GPU_SRCS "$ROOT/tests/cpp/gtest_shim/gtest_main.cpp"; do
  o="$OBJ/$(basename "${src%.cpp}").o"
  $CXX $FLAGS -c "$src" -o "$o"
  objs="$objs $o"
done
build_test() {  # name, source
  $CXX $FLAGS "$2" $objs $LINK -o "$OUT/$1"
  echo "built oracle/_ref/bin/$1  <-  ${2#$FARM/}"
}
# the reference's GPU acceptance tests (CPU == GPU) and GPU benchmark printers
build_test test_field_arithmetic_cuda "$S/poseidon/test/test_field_arithmetic_cuda.cpp"
build_test test_poseidon_cuda         "$S/poseidon/test/test_poseidon_cuda.cpp"
build_test test_merkle_tree_cuda      "$S/merkle_tree/test/test_merkle_tree_cuda.cpp"
build_test test_merkle_benchmark_cuda "$S/merkle_tree/test/test_merkle_benchmark_cuda.cpp"
# the reference's CPU suites against its own CPU code: a sanity gate for the oracle and for the gtest shim
build_test test_field_arithmetic      "$S/poseidon/test/test_field_arithmetic.cpp"
build_test test_field_accumulation    "$S/poseidon/test/test_field_accumulation.cpp"
build_test test_poseidon              "$S/poseidon/test/test_poseidon.cpp"
build_test test_merkle_tree           "$S/merkle_tree/test/test_merkle_tree.cpp"
# the reference's benchmark driver (has its own main): poseidon_benchmark, CUDA_ENABLED as src/poseidon/CMakeLists.txt:40-43 sets it
nomain=$(echo $objs | tr ' ' '\n' | grep -v gtest_main | tr '\n' ' ')
$CXX $FLAGS -DCUDA_ENABLED -I"${CUDA_HOME:-/usr/local/cuda}/include" "$S/poseidon/test/benchmark.cpp" $nomain $LINK -o "$OUT/poseidon_benchmark"
echo "built oracle/_ref/bin/poseidon_benchmark  <-  src/poseidon/test/benchmark.cpp"
# the reference's profiling driver (a loop over batch sizes for an external profiler; its own main)
$CXX $FLAGS "$S/poseidon/cuda/poseidon_cuda_profiler.cpp" $nomain $LINK -o "$OUT/poseidon_cuda_profiler"
echo "built oracle/_ref/bin/poseidon_cuda_profiler  <-  src/poseidon/cuda/poseidon_cuda_profiler.cpp"
rm -rf "$FARM"

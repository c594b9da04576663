#!/bin/bash
# TEST INFRASTRUCTURE: builds tests/emu/_build/libgenfft_emu.so -- genfft_b200/csrc/*.cu compiled by g++ against the
# fake CUDA runtime in shim/ (see shim/cuda_runtime.h).  Used only by tests/test_emu_*.py; never loaded by genfft_b200.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/../../genfft_b200/csrc"
OUT="${GENFFT_EMU_OUT:-$HERE/_build}"
OBJ="$OUT/obj"
mkdir -p "$OBJ"
rm -f "$OBJ/plan.o"
CXX=${CXX:-g++}
FLAGS="-std=c++17 ${GENFFT_EMU_OPT:--O1} $GENFFT_EMU_EXTRA -fPIC -DGENFFT_EMU=1 -I$HERE/shim -Wno-unknown-pragmas -Wno-attributes -x c++"
pids=()
for k in 0 1 2 3; do
  $CXX $FLAGS -DGENFFT_CSET=$k -c "$SRC/chains_inst.cu" -o "$OBJ/chains_$k.o" & pids+=($!)
done
for k in 0 1 2 3 4 5; do
  $CXX $FLAGS -DGENFFT_KSET=$k -c "$SRC/kernels_inst.cu" -o "$OBJ/kernels_$k.o" & pids+=($!)
done
for f in planner pass_chain abi; do
  $CXX $FLAGS -c "$SRC/$f.cu" -o "$OBJ/$f.o" & pids+=($!)
done
$CXX $FLAGS -c "$SRC/host_exec.cu" -o "$OBJ/host_exec.o" & pids+=($!)
$CXX $FLAGS -c "$HERE/emu_runtime.cpp" -o "$OBJ/emu_runtime.o" & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
$CXX -shared -o "$OUT/libgenfft_emu.so" "$OBJ"/*.o -lpthread
echo "built $OUT/libgenfft_emu.so"

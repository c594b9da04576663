#!/bin/bash
# Builds genfft_b200/lib/libgenfft_cuda.so for sm_100a (in-tree, so the .so travels to the GPU box).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${GENFFT_LIB_OUT:-$HERE/../lib}"  # GENFFT_LIB_OUT + GENFFT_NVCC_EXTRA: variant builds for A/B measurements
OBJ="$OUT/obj"
mkdir -p "$OUT" "$OBJ"
rm -f "$OBJ/plan.o"
NVCC=${NVCC:-nvcc}
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $GENFFT_NVCC_EXTRA"
pids=()
for k in 0 1 2 3; do
  $NVCC $FLAGS -DGENFFT_CSET=$k -c "$HERE/chains_inst.cu" -o "$OBJ/chains_$k.o" & pids+=($!)
done
for k in 0 1 2 3 4 5; do
  $NVCC $FLAGS -DGENFFT_KSET=$k -c "$HERE/kernels_inst.cu" -o "$OBJ/kernels_$k.o" & pids+=($!)
done
for f in planner pass_chain abi; do
  $NVCC $FLAGS -c "$HERE/$f.cu" -o "$OBJ/$f.o" & pids+=($!)
done
$NVCC $FLAGS -c "$HERE/host_exec.cu" -o "$OBJ/host_exec.o" & pids+=($!)
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libgenfft_cuda.so" "$OBJ"/kernels_*.o "$OBJ"/chains_*.o "$OBJ/planner.o" "$OBJ/pass_chain.o" "$OBJ/abi.o" "$OBJ/host_exec.o"
echo "built $OUT/libgenfft_cuda.so"

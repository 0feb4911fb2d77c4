#!/bin/bash
# Build libcsam_sm100.so in-tree (sm_100a only).  Usage: build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${CSAM_BUILD_OUT:-$HERE/../_C}"     # CSAM_BUILD_OUT: side builds (e.g. -DCSAM_TRACE) next to the product library
mkdir -p "$OUT"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
pids=()
for f in abi gemm gemm_pair elementwise attention_simt attention_tc decoder_fused post nms cc; do
  $NVCC $FLAGS "$@" -c "$HERE/$f.cu" -o "$OUT/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libcsam_sm100.so" "$OUT"/*.o -cudart static
echo "built $OUT/libcsam_sm100.so"

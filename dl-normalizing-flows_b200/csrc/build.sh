#!/usr/bin/env bash
# Build librnvp_b200.so (sm_100a only) in-tree.  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../librnvp_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
       -Xcompiler -Wall -Xcompiler -Wno-unused-function --expt-relaxed-constexpr "$@")
mkdir -p "$HERE/build"
pids=()
for f in elementwise conv_simt conv_tc runtime dp optim; do
  "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/build/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$HERE"/build/{elementwise,conv_simt,conv_tc,runtime,dp,optim}.o -lcudart -lcuda -ldl
echo "built $OUT"

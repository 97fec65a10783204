#!/bin/bash
# Build libbsplinekit_b200.so for sm_100a, in-tree (the .so travels to the GPU box).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/lib"
OBJ="$HERE/build"
mkdir -p "$OUT" "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v"
pids=()
for f in bsk_api bsk_eval bsk_solve bsk_host bsk_eval_cell bsk_eval_cell_f32 bsk_solve_win bsk_multi bsk_cyclic_banded bsk_galerkin bsk_solve_f32 bsk_probe; do
  src="$HERE/csrc/$f.cu"
  extra=""
  if [ "$f" = "bsk_eval_cell_f32" ]; then src="$HERE/csrc/bsk_eval_cell.cu"; extra="-DBSK_CELL_F32"; fi
  obj="$OBJ/$f.o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$HERE/csrc/core.cuh" -nt "$obj" ] || [ "$HERE/csrc/handles.cuh" -nt "$obj" ] || [ "$HERE/csrc/pp.cuh" -nt "$obj" ] || [ "$HERE/csrc/tma.cuh" -nt "$obj" ] || [ "$HERE/csrc/banfac.cuh" -nt "$obj" ] || [ "$HERE/csrc/banslv.cuh" -nt "$obj" ] || [ "$HERE/csrc/pptable.h" -nt "$obj" ] || [ "$HERE/csrc/dd.h" -nt "$obj" ] || [ "$HERE/../include/bsk.h" -nt "$obj" ]; then
    ( $NVCC $FLAGS $extra -c "$src" -o "$obj" > "$OBJ/$f.log" 2>&1 || { cat "$OBJ/$f.log" | grep -v "^ptxas info" | head -60; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o "$OUT/libbsplinekit_b200.so" "$OBJ"/bsk_api.o "$OBJ"/bsk_eval.o "$OBJ"/bsk_solve.o "$OBJ"/bsk_host.o "$OBJ"/bsk_eval_cell.o "$OBJ"/bsk_eval_cell_f32.o "$OBJ"/bsk_solve_win.o "$OBJ"/bsk_multi.o "$OBJ"/bsk_cyclic_banded.o "$OBJ"/bsk_galerkin.o "$OBJ"/bsk_solve_f32.o "$OBJ"/bsk_probe.o -lcudart_static -ldl -lrt -lpthread
echo "built $OUT/libbsplinekit_b200.so"

#!/bin/bash
# A/B of the side-entry (gather-bound) workloads between a previous build and the in-tree library, plus a sweep of the
# x-panel width on one config-5 row block: tools/ab_side.sh <prev.so> <out.log>
PREV=$1; OUT=$2; : > $OUT
run() {
  for lib in prev new; do
    if [ $lib = prev ]; then export TILESPMV_LIB_PATH=$PREV; else unset TILESPMV_LIB_PATH; fi
    echo "== $lib: $*" >> $OUT
    python tools/spmv_run.py "$@" 2>&1 | grep -v Warning | grep -v "torch.sparse_csr" | grep -v "^  A = " >> $OUT
  done
}
run --workload uniform --n 1048576 --iters 100 --check
run --workload rmat --scale 20 --iters 100 --check
run --workload rmat --scale 22 --precision f32 --iters 50
run --workload lap3d27 --grid 160 --iters 200
unset TILESPMV_LIB_PATH
for pb in 0 41943040 33554432 25165824 16777216; do
  echo "== new: c5 shard xpanel_bytes=$pb" >> $OUT
  python tools/spmv_run.py --workload uniform --n 50000000 --rows 6250000 --iters 30 --xpanel-bytes $pb 2>&1 | grep "us/SpMV" >> $OUT
done

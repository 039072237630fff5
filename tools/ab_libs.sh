#!/bin/bash
# A/B of two builds of libtilespmv_b200.so on ONE box: tools/ab_libs.sh <prev.so> <out.log>
# (the current in-tree library is "new"); every workload runs prev, new, prev, new
PREV=$1; OUT=$2; : > $OUT
run() { # label, args...
  for rep in 1 2; do
    for lib in prev new; do
      if [ $lib = prev ]; then export TILESPMV_LIB_PATH=$PREV; else unset TILESPMV_LIB_PATH; fi
      echo "== $lib rep$rep: $*" >> $OUT
      python tools/spmv_run.py "$@" 2>&1 | grep -v Warning | grep -v "torch.sparse_csr" | grep -v "^  A = " >> $OUT
    done
  done
}
run --workload uniform --n 1048576 --iters 100 --check
run --workload uniform --n 50000000 --rows 6250000 --iters 30
run --workload rmat --scale 20 --iters 100 --check
run --workload rmat --scale 22 --precision f32 --iters 50
run --workload lap3d27 --grid 160 --iters 200
run --workload lap2d --grid 1024 --iters 300 --check
run --workload banded --n 1048576 --iters 200

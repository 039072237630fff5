#!/bin/bash
# A/B of the specialised (plain) epilogue on ONE box: tools/ab_epilogue.sh <out.log>
# TILESPMV_NO_PLAIN_EPILOGUE=1 routes every launch through the general epilogue (the round-2 baseline); every workload
# runs general, plain, general, plain
OUT=$1; : > $OUT
run() {
  for rep in 1 2; do
    for mode in general plain; do
      if [ $mode = general ]; then export TILESPMV_NO_PLAIN_EPILOGUE=1; else unset TILESPMV_NO_PLAIN_EPILOGUE; fi
      echo "== $mode rep$rep: $*" >> $OUT
      python tools/spmv_run.py "$@" 2>&1 | grep -v Warning | grep -v "torch.sparse_csr" | grep -v "^  A = " >> $OUT
    done
  done
  unset TILESPMV_NO_PLAIN_EPILOGUE
}
run --workload lap3d27 --grid 160 --iters 300
run --workload lap2d --grid 1024 --iters 500 --check
run --workload banded --n 1048576 --iters 300
run --workload uniform --n 1048576 --iters 100 --check
run --workload rmat --scale 20 --iters 100 --check

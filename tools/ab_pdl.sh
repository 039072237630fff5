#!/bin/bash
# A/B of programmatic dependent launch on ONE box: tools/ab_pdl.sh <out.log>
# TILESPMV_NO_PDL=1 launches every kernel with the full stream order (the baseline)
OUT=$1; : > $OUT
run() {
  reps=$1; shift
  for rep in $(seq $reps); do
    for mode in nopdl pdl; do
      if [ $mode = nopdl ]; then export TILESPMV_NO_PDL=1; else unset TILESPMV_NO_PDL; fi
      echo "== $mode rep$rep: $*" >> $OUT
      python tools/spmv_run.py "$@" 2>&1 | grep -v Warning | grep -v "torch.sparse_csr" | grep -v "^  A = " >> $OUT
    done
  done
  unset TILESPMV_NO_PDL
}
run 2 --workload lap3d27 --grid 160 --iters 300
run 2 --workload lap2d --grid 1024 --iters 500 --check
run 2 --workload banded --n 1048576 --iters 300
run 1 --workload uniform --n 1048576 --iters 100 --check
run 1 --workload uniform --n 8000000 --rows 1000000 --iters 50 --xpanel-bytes 16000000
echo "== iterate probe, TILESPMV_NO_PDL_IN_GRAPHS=1 (graphs without the programmatic edge)" >> $OUT
TILESPMV_NO_PDL_IN_GRAPHS=1 python tools/iterate_probe.py --iters 400 >> $OUT 2>&1
echo "== iterate probe, default (programmatic edges inside the graph)" >> $OUT
python tools/iterate_probe.py --iters 400 >> $OUT 2>&1

# A/B: split fix-up kernels as links of the programmatic-launch chain (default) vs plain launches
OUT=$1; : > $OUT
run() { for mode in plain chain plain chain; do
  if [ $mode = plain ]; then export TILESPMV_NO_PDL_FIXUP=1; else unset TILESPMV_NO_PDL_FIXUP; fi
  echo "== $mode: $*" >> $OUT
  python tools/spmv_run.py "$@" 2>&1 | grep -v Warning | grep -v "torch.sparse_csr" | grep -v "^  A = " | sed -e "s/.*| gen/gen/" >> $OUT
done; unset TILESPMV_NO_PDL_FIXUP; }
run --workload rmat --scale 20 --iters 200 --check
run --workload rmat --scale 22 --precision f32 --iters 60

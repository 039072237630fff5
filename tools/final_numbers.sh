# table rows of DESIGN.md section 4 with the final build (one B200)
run() { python tools/spmv_run.py "$@" 2>&1 | grep -v Warning | grep -v "torch.sparse_csr" | grep -v "^  A = " | sed -e "s/.*| gen/gen/"; }
echo "== rmat 20 f64"; run --workload rmat --scale 20 --iters 100 --check
echo "== rmat 22 f32"; run --workload rmat --scale 22 --precision f32 --iters 50 --check
echo "== config-5 row block (uniform 6.25 M x 50 M)"; run --workload uniform --n 50000000 --rows 6250000 --iters 30
echo "== band_contig 1 M"; run --workload band_contig --n 1048576 --iters 300
echo "== rmat 24 f32 (config 4)"; run --workload rmat --scale 24 --precision f32 --iters 20 --check

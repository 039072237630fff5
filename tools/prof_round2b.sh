# final 1-GPU evidence of round 2 (after the specialised epilogue + programmatic dependent launch): GPU suite, smoke,
# bench lines of both arms, ncu launch list + three --set full captures (each program first exits 0 without ncu), warp sweep
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_suite6.log 2>&1; tail -3 gpurun_out/r2_gpu_suite6.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_1gpu_final2.json 2> gpurun_out/r2_bench_1gpu_final2.err; cut -c1-300 gpurun_out/r2_bench_1gpu_final2.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_1gpu_20steps.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_1gpu_20steps.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref_1gpu2.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_ref_1gpu2.json
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
NCU="ncu --set full --clock-control none --import-source on -k regex:tile_spmv"
$B > gpurun_out/r2_b5b.json 2> gpurun_out/r2_b5b.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_bench_steps5.csv $B > gpurun_out/ncu_launch_b.log 2>&1
$NCU -s 5 -c 1 -o gpurun_out/prof_r2b_bench $B > gpurun_out/ncu_r2b_a.log 2>&1
U="python tools/spmv_run.py --workload uniform --n 1048576 --iters 2 --warmup 2"
$U > /dev/null 2>&1 && $NCU -s 3 -c 1 -o gpurun_out/prof_r2b_uniform $U > gpurun_out/ncu_r2b_b.log 2>&1
L="python tools/spmv_run.py --workload lap2d --grid 1024 --iters 2 --warmup 2"
$L > /dev/null 2>&1 && $NCU -s 3 -c 1 -o gpurun_out/prof_r2b_lap2d $L > gpurun_out/ncu_r2b_c.log 2>&1
ls -la gpurun_out/prof_r2b*.ncu-rep
python tools/ab.py --workload lap3d27 --grid 160 --configs "2:0,2:21,2:18,2:16" --rounds 2 --iters 200 > gpurun_out/r2b_warp_sweep.log 2>&1; cat gpurun_out/r2b_warp_sweep.log

set -x
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
NCU="ncu --set full --clock-control none --import-source on -k regex:tile_spmv"
$B > gpurun_out/r2_b5.json 2> gpurun_out/r2_b5.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps5.csv $B > gpurun_out/ncu_launch.log 2>&1
$NCU -s 5 -c 1 -o gpurun_out/prof_r2_bench $B > gpurun_out/ncu_r2a.log 2>&1
U="python tools/spmv_run.py --workload uniform --n 1048576 --iters 2 --warmup 2"
$U > /dev/null 2>&1 && $NCU -s 3 -c 1 -o gpurun_out/prof_r2_uniform $U > gpurun_out/ncu_r2b.log 2>&1
C="python tools/spmv_run.py --workload uniform --n 50000000 --rows 6250000 --iters 2 --warmup 1"
$C > /dev/null 2>&1 && $NCU -s 10 -c 1 -o gpurun_out/prof_r2_c5panel $C > gpurun_out/ncu_r2c.log 2>&1
L="python tools/spmv_run.py --workload lap2d --grid 1024 --iters 2 --warmup 2"
$L > /dev/null 2>&1 && $NCU -s 3 -c 1 -o gpurun_out/prof_r2_lap2d $L > gpurun_out/ncu_r2d.log 2>&1
R="python tools/spmv_run.py --workload rmat --scale 20 --iters 2 --warmup 2"
$R > /dev/null 2>&1 && $NCU -s 3 -c 1 -o gpurun_out/prof_r2_rmat20 $R > gpurun_out/ncu_r2e.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -6
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_suite3.log 2>&1; tail -3 gpurun_out/r2_gpu_suite3.log
python bench.py > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err; cut -c1-300 gpurun_out/r2_bench_1gpu_final.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref_1gpu.json 2>/dev/null; cut -c1-200 gpurun_out/r2_bench_ref_1gpu.json

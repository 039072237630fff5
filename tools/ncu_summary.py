"""Summarise an .ncu-rep: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [chunks_per_launch]
Prints the headline metrics and the per-chunk instruction profile of the SASS (hot lines)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
nchunks = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio']
for r in rows[2:]:
    print("KERNEL", r[h.index('Kernel Name')][:90])
    for k in want:
        if k in h:
            print(f"  {k:90s} {r[h.index(k)]} {u[h.index(k)]}")
if nchunks:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    iS, iI, iSrc = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed'), hdr.index('Source')
    iW, iWi = hdr.index('L1 Wavefronts Shared'), hdr.index('L1 Wavefronts Shared Ideal')
    data = []
    for r in rows[2:]:
        if r and r[0] == 'Kernel Name':  # second kernel of the report: only the first one is listed
            break
        if len(r) > max(iS, iI) and r[iS] != '':
            data.append(r)
    tot_s = sum(int(r[iS]) for r in data)
    tot_i = sum(int(r[iI]) for r in data)
    print(f"total samples {tot_s}, warp instructions {tot_i} = {tot_i / nchunks:.1f} per chunk")
    for n, r in enumerate(data):
        i = int(r[iI])
        if i / nchunks > 0.3:
            print(f"{n:5d} {i / nchunks:7.2f} {int(r[iS]):5d} {r[iW]:>9} {r[iWi]:>9}  {r[iSrc].strip()[:80]}")

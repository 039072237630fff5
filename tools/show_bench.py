"""Pretty-print the multi-GPU keys of bench.py JSON lines: python tools/show_bench.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "ERR", e)
        continue
    print("=====", f)
    print("  value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), d["config"].get("headline_exchange"), "| steps", d["steps"])
    if "iterate" in d:
        for k in ("nccl", "fused", "pipelined", "halo"):
            it = d["iterate"].get(k, {})
            print("   ", k, it.get("ms_per_iteration"), it.get("gpu_launches"), d["iterate"]["verification"].get(k, {}).get("verified"),
                  d["iterate"]["verification"].get(k, {}).get("max_err_over_bound"), d["iterate"]["verification"].get(k, {}).get("error"))
        print("    spmv_no_exchange", d["spmv_no_exchange"]["ms_per_step"])
        if d.get("one_gpu"):
            print("    one_gpu", d["one_gpu"]["ms_per_iteration"], "build_s", round(d["one_gpu"]["build_s"], 1), "speedup", d["speedup_vs_one_gpu"])
        print("    weak", d.get("weak_scaling_stencil"))
        if d.get("e2e"):
            print("    e2e", round(d["e2e"]["value"], 1), d["e2e"]["ms_per_step"])
        print("    roofline frac", round(d["roofline"]["frac"], 3), "units", d["extra"]["launch_units"], d["extra"]["unit_deps_rank0"],
              "gen", round(d["extra"]["gen_s"], 1), "setup", round(d["extra"]["convert_plan_s"], 1), "clocks", d.get("clocks"))

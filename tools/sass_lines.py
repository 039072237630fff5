"""Static SASS of one kernel of libtilespmv_b200.so, grouped by the source line it was generated from (needs the
-lineinfo build): python tools/sass_lines.py <substring of the mangled kernel name> [first_line last_line] [--count]

  --count   only the number of SASS instructions per source line (no listing)

Used to check the instruction budget of a code path on the CPU box before spending GPU time on it (the per-line
EXECUTED counts come from an ncu capture: ncu -i rep --page source --csv --print-source cuda,sass)."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("TILESPMV_LIB_PATH", os.path.join(ROOT, "tilespmv_b200", "libtilespmv_b200.so"))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    count_only = "--count" in sys.argv
    name = args[0]
    lo, hi = (int(args[1]), int(args[2])) if len(args) >= 3 else (0, 1 << 30)
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "spmv.sm_100a.cubin", LIB], cwd=d, capture_output=True)
        cubins = [f for f in os.listdir(d) if f.endswith(".cubin")]
        if not cubins:
            raise SystemExit("no spmv cubin in " + LIB)
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, cubins[0])], capture_output=True, text=True).stdout
    sect, line, per, order = None, None, {}, []
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            sect = m.group(1)
            continue
        if sect is None or name not in sect:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m and line:
            if line not in per:
                per[line] = []
                order.append(line)
            per[line].append(m.group(1).strip())
    total = 0
    for key in sorted(per):
        if key[0] != "spmv.cu" or not (lo <= key[1] <= hi):
            continue
        total += len(per[key])
        if count_only:
            print(f"{key[1]:5d} {len(per[key]):4d}")
        else:
            print(f"--- {key[0]}:{key[1]} ({len(per[key])})")
            for i in per[key]:
                print("      " + i)
    print(f"total {total} SASS instructions in spmv.cu lines {lo}..{hi} of *{name}*")


if __name__ == "__main__":
    main()

"""Summarise an `ncu --set full` report into profiles/: a markdown table per kernel and the per-launch DRAM traffic
JSON that bench.py reads for `roofline.traffic`.

    python tools/ncu_summary.py gpurun_out/prof_conv.ncu-rep profiles/r01_ncu_conv_summary.md profiles/r01_traffic.json [title] [frames]
"""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(rep, md_out, json_out, title, frames):
    if rep.endswith(".csv"):      # the raw page exported on the GPU box (reports above 64 MiB do not travel back)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    out = [f"# {title}", "",
           f"Workload: `python tools/r02_kernels.py` (default plan: {frames} frames of 752x480 per launch; exact-mode plan: 16 frames; matcher "
           "2001 x 1777 rows; guided search m = 1000). ncu replays each kernel ~40x cold-cache and serialised: compare shares and ratios, "
           "not absolute times.", ""]
    traffic = {"frames_per_launch": frames}
    for r in rows[2:]:
        name = r[kcol]
        out += [f"### `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
        by = 0.0
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                out.append(f"| {m} | {r[i]} | {units[i]} |")
                if m.startswith("dram__bytes"):
                    by += float(r[i].replace(",", "")) * TO_BYTES.get(units[i], 1.0)
        out.append("")
        key = re.sub(r"^void ", "", name)
        key = key[:key.rindex("(")] if "(" in key else key      # drop the parameter list, keep template arguments
        traffic.setdefault(key, by)
        if "sm__throughput.avg.pct_of_peak_sustained_elapsed" in hdr:
            out.insert(len(out) - 1, f"| sm__throughput.avg.pct_of_peak_sustained_elapsed | {r[hdr.index('sm__throughput.avg.pct_of_peak_sustained_elapsed')]} | % |")
    open(md_out, "w").write("\n".join(out))
    json.dump(traffic, open(json_out, "w"), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu --set full (--clock-control none)",
         int(sys.argv[5]) if len(sys.argv) > 5 else 64)

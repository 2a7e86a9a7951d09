"""Turns the raw ncu output of tools/gpu_profile.sh (gpurun_out/) into the text summaries kept under profiles/.

  python tools/summarize_profiles.py r01        # writes profiles/r01_launches_config2_summary.txt, r01_ncu_full_*.txt
"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEEP = [
    "dram__bytes_read.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size", "launch__grid_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.avg",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__sass_l1tex_t_requests_pipe_lsu_mem_global_op_ldgsts.sum",
    "sm__sass_l1tex_t_sectors_pipe_lsu_mem_global_op_ldgsts_cache_bypass.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
]


def launches(tag):
    path = os.path.join(OUT, "launches_full.csv")
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    n = 0
    for row in r:
        try:
            ns = float(row[vi].replace(",", ""))
        except (ValueError, IndexError):
            continue
        name = row[ki].split("(")[0].replace("void ", "").split("<")[0].replace("rbl::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        n += 1
    tot = sum(v[1] for v in agg.values())
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none : one whole config-2 solve (python tools/profile_solve.py)",
           "# serialised, cold-cache launches: compare SHARES with bench.py's CUDA-event numbers, not absolutes",
           f"total device time {tot / 1e9:.3f} s over {n} launches"]
    for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{name:<30} launches={cnt:5d}  total={ns / 1e9:.4f} s  share={100 * ns / tot:5.1f}%  avg={ns / cnt / 1e3:.1f} us")
    with open(os.path.join(PROF, f"{tag}_launches_config2_summary.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out[:12]))


def full(tag, rep, name):
    path = os.path.join(OUT, rep)
    if not os.path.exists(path):
        print("missing", path)
        return
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = []
    for h, u, v in zip(hdr, units, vals):
        if h == "Kernel Name" or h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            out.append(f"{h:<90} {v} {u}")
    with open(os.path.join(PROF, f"{tag}_ncu_full_{name}.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    print(name, [l for l in out if "dram__bytes_read.sum.per_second" in l or "gpu__time_duration" in l])


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    launches(tag)
    full(tag, "prof_gram_h.ncu-rep", "presplit_gram_h")
    full(tag, "prof_update_h.ncu-rep", "presplit_update_h")
    full(tag, "prof_ritz_h.ncu-rep", "ritz_h")

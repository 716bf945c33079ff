"""Evidence post-processing (run in the build container on what a gpurun call brought back):
  summarise_ncu.py shares <launch list csv> <out json>      per-kernel shares of one bench step
  summarise_ncu.py full <out csv> <rep> [<rep> ...]          key metrics of ncu --set full captures (ncu -i --page raw)"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def rows_of(text):
    lines = [ln for ln in text.splitlines() if ln.startswith('"')]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def shares(src, dst):
    rows = rows_of(open(src).read())
    head = rows[0]
    name_i, metric_i, value_i, unit_i = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value"), head.index("Metric Unit")
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if r[metric_i] != "gpu__time_duration.sum":
            continue
        v = float(r[value_i].replace(",", ""))
        ms = v / 1e6 if r[unit_i] in ("ns", "nsecond") else v / 1e3 if r[unit_i] in ("us", "usecond") else v
        nm = r[name_i].replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        base = re.sub(r"<.*", "", re.sub(r"\(.*", "", nm)).split("::")[-1].strip()
        if base == "conv_gemm_kernel":  # split by engine: the DAC decoder's instances use the LeakyReLU / Snake epilogues
            m = re.search(r"conv_gemm_kernel<\(int\)(-?\d+), \(int\)(-?\d+), \(int\)(-?\d+)", nm) or re.search(r"conv_gemm_kernel<(-?\d+), (-?\d+), (-?\d+)", nm)
            if m:
                base += " (DAC)" if (m.group(1) in ("1", "4") or m.group(3) == "3") else " (estimator)"
        a = agg.setdefault(base, {"launches": 0, "ms": 0.0})
        a["launches"] += 1
        a["ms"] += ms
        total += ms
    for a in agg.values():
        a["share"] = a["ms"] / total
    out = {"source": f"{src} (ncu --metrics gpu__time_duration.sum --clock-control none, one timed bench step, LS_NCU_RANGE=1)",
           "total_ms": total, "kernels": dict(sorted(agg.items(), key=lambda kv: -kv[1]["ms"]))}
    json.dump(out, open(dst, "w"), indent=1)
    for k, a in list(out["kernels"].items())[:6]:
        print(f"{k:32s} {a['launches']:5d} launches {a['ms']:8.3f} ms  {100 * a['share']:5.1f} %")
    print("total", round(total, 3), "ms")


def full(dst, reps):
    out_rows = []
    for rep in reps:
        text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = rows_of(text)
        head, units = rows[0], rows[1]
        for r in rows[2:]:
            rec = {"capture": rep.split("/")[-1], "Kernel Name": r[head.index("Kernel Name")][:120]}
            for k in KEYS:
                if k in head:
                    i = head.index(k)
                    rec[k] = f"{r[i]} {units[i]}".strip()
            out_rows.append(rec)
    cols = ["capture", "Kernel Name"] + [k for k in KEYS if any(k in r for r in out_rows)]
    with open(dst, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for r in out_rows:
            w.writerow(r)
    for r in out_rows:
        print(r["capture"], r.get("gpu__time_duration.sum"), "read", r.get("dram__bytes_read.sum"), "write", r.get("dram__bytes_write.sum"),
              "dram%", r.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "l1tex%", r.get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"))


if __name__ == "__main__":
    if sys.argv[1] == "shares":
        shares(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3:])

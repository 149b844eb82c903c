"""Turns `ncu --set full` reports brought back in gpurun_out/ into the committed evidence under profiles/:
  profiles/<name>_raw.csv      the metrics the roofline arithmetic and DESIGN.md quote, one column per captured launch
  profiles/<name>_summary.md   the same as a table + derived numbers
  profiles/ncu_traffic.json    dram bytes per launch keyed by workload (read by bench.py for roofline.traffic)
usage: python scripts/ncu_summary.py <report.ncu-rep> <name> [workload pairs_per_launch[,pairs_per_launch...]]
(runs `ncu -i ... --page raw --csv` here; no GPU needed)"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(
    r"^(Kernel Name|gpu__time_duration\.sum|dram__bytes_read\.sum|dram__bytes_write\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"launch__grid_size|launch__block_size|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|lts__t_sector_hit_rate\.pct|"
    r"sass__inst_executed_local_loads|sass__inst_executed_local_stores|smsp__inst_executed\.sum|smsp__thread_inst_executed\.sum|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_(fma|alu|xu|lsu)\.avg\.pct_of_peak_sustained_active|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed)$")


def main():
    rep, name = sys.argv[1], sys.argv[2]
    workload = sys.argv[3] if len(sys.argv) > 3 else None
    ppl = [float(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else []
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if KEEP.match(h)]
    out_csv = os.path.join(ROOT, "profiles", name + "_raw.csv")
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch %d" % i for i in range(len(data))])
        for i in cols:
            w.writerow([hdr[i], units[i]] + [r[i] for r in data])
    H = {h: i for i, h in enumerate(hdr)}

    def val(r, m):
        try:
            return float(r[H[m]])
        except Exception:
            return float("nan")

    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    lines = ["# %s — `ncu --set full --clock-control none`, report %s" % (name, os.path.basename(rep)), "",
             "| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |", "|---|---|" + "---|" * len(data)]
    for i in cols:
        lines.append("| %s | %s | %s |" % (hdr[i], units[i], " | ".join(r[i] for r in data)))
    lines.append("")
    traffic = []
    for k, r in enumerate(data):
        rd = val(r, "dram__bytes_read.sum") * scale.get(units[H["dram__bytes_read.sum"]], 1)
        wr = val(r, "dram__bytes_write.sum") * scale.get(units[H["dram__bytes_write.sum"]], 1)
        du = val(r, "gpu__time_duration.sum")
        du_s = du * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(units[H["gpu__time_duration.sum"]], 1e-3)
        extra = ""
        if k < len(ppl):
            extra = " = %.3f GB per pair (%g pairs in this launch)" % ((rd + wr) / ppl[k] / 1e9, ppl[k])
            traffic.append({"pairs_per_launch": ppl[k], "dram_bytes_per_launch": rd + wr, "duration_ms": du_s * 1e3,
                            "source": "profiles/%s_raw.csv launch %d" % (name, k)})
        lines.append("launch %d: DRAM read + write %.3f GB%s; duration %.3f ms -> %.0f GB/s of DRAM traffic" % (
            k, (rd + wr) / 1e9, extra, du_s * 1e3, (rd + wr) / du_s / 1e9))
    open(os.path.join(ROOT, "profiles", name + "_summary.md"), "w").write("\n".join(lines) + "\n")
    if workload and traffic:
        p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        j = json.load(open(p)) if os.path.exists(p) else {}
        j[workload] = traffic
        json.dump(j, open(p, "w"), indent=1)
    print("\n".join(lines[-len(data) - 1:]))


if __name__ == "__main__":
    main()

"""Summarises an ncu report of the step kernels into profiles/ (text) and records their DRAM traffic per step in
profiles/step_traffic.json, which bench.py reads at run time for roofline.traffic.

usage: python tools/ncu_summary.py report.ncu-rep <traffic key, e.g. wiki6b/replay/B65536> <out.txt> [launches.csv out.md]"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]


def main():
    rep, key, out_txt = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    iname = hdr.index("Kernel Name")
    by = {}
    for r in body:
        by.setdefault(r[iname], []).append(r)
    lines = ["ncu --set full --clock-control none --cache-control none --import-source on  (%s; averages over the captured launches)"
             % os.path.basename(rep), ""]
    traffic = 0.0
    for name, rs in by.items():
        lines.append("== %s   (%d launches)" % (name, len(rs)))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals = [float(r[i].replace(",", "")) for r in rs if r[i] not in ("", "n/a")]
                if vals:
                    lines.append("   %-72s %14.4f %s" % (k, sum(vals) / len(vals), units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "_not_issued" not in h:
                vals = [float(r[i]) for r in rs if r[i] not in ("", "n/a")]
                if vals and sum(vals) / len(vals) > 0.05:
                    stalls.append((sum(vals) / len(vals), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        lines.append("   warp stall cycles per issue: " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)))
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[rd]]
        per = sum(float(r[rd]) + float(r[wr]) for r in rs) / len(rs) * scale
        lines.append("   dram bytes per launch (read + write): %.1f MB" % (per / 1e6))
        traffic += per
        lines.append("")
    lines.append("DRAM traffic of one step (sum over the kernels above): %.1f MB" % (traffic / 1e6))
    open(out_txt, "w").write("\n".join(lines) + "\n")
    path = os.path.join(ROOT, "profiles", "step_traffic.json")
    try:
        t = json.load(open(path))
    except Exception:
        t = {}
    t[key] = {"dram_bytes": traffic, "source": os.path.relpath(out_txt, ROOT)}
    json.dump(t, open(path, "w"), indent=1, sort_keys=True)
    print("\n".join(lines[-12:]))
    if len(sys.argv) > 5:   # launch list (gpu__time_duration.sum per launch) -> share of each kernel in the step
        rows = list(csv.reader(open(sys.argv[4])))
        start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        h = rows[start]
        ik, iv, im = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
        acc = {}
        for r in rows[start + 1:]:
            if len(r) > iv and r[im] == "gpu__time_duration.sum":
                a = acc.setdefault(r[ik].split("(")[0], [0, 0.0])
                a[0] += 1; a[1] += float(r[iv].replace(",", ""))
        tot = sum(v[1] for v in acc.values())
        md = ["| kernel | launches | total us | share |", "|---|---|---|---|"]
        for k, (n, ns) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
            md.append("| `%s` | %d | %.1f | %.1f %% |" % (k, n, ns / 1e3, 100 * ns / tot))
        open(sys.argv[5], "w").write("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: shares, not absolutes)\n\n" + "\n".join(md) + "\n")


if __name__ == "__main__":
    main()

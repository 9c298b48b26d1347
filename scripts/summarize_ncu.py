"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.
  python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md "<command>"
  python scripts/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r1_propagate_ncu.md "<command>"
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    h = {k: i for i, k in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*$", "", r[h["Kernel Name"]]).replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, r[h["Grid Size"]], r[h["Block Size"]]])
        a[0] += 1
        a[1] += float(r[h["Metric Value"]]) / 1e3
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("command: `%s`\n\nPer-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n" % cmd)
        f.write("| kernel | launches | total us | avg us | share | grid | block |\n|---|---:|---:|---:|---:|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f%% | %s | %s |\n" % (name, a[0], a[1], a[1] / a[0], 100 * a[1] / total, a[2], a[3]))
        f.write("\ntotal kernel time %.1f us over %d launches\n" % (total, len(rows)))


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
           "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]


def full(src, dst, cmd):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    h = {k: i for i, k in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full summary\n\ncommand: `%s`\n\n" % cmd)
        for r in rows:
            f.write("## `%s`  grid %s block %s\n\n| metric | value | unit |\n|---|---:|---|\n" % (
                r[h["Kernel Name"]], r[h.get("Grid Size", 0)], r[h.get("Block Size", 0)]))
            for m in METRICS:
                if m in h:
                    f.write("| %s | %s | %s |\n" % (m, r[h[m]], units[h[m]]))
            if "dram__bytes_read.sum" in h:
                def tobytes(v, u):
                    v = float(v)
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                tr = tobytes(r[h["dram__bytes_read.sum"]], units[h["dram__bytes_read.sum"]]) + tobytes(r[h["dram__bytes_write.sum"]], units[h["dram__bytes_write.sum"]])
                f.write("\nDRAM traffic per launch: %.1f MB\n\n" % (tr / 1e6))


def traffic(src, dst, cmd):
    """profiles/traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every hot kernel's
    full-size launches in an `ncu --set full` capture (the early-exit launches of non-resampling steps are skipped)."""
    import json
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    h = {k: i for i, k in enumerate(hdr)}
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def val(r, k):
        return float(r[h[k]]) * mult.get(units[h[k]], 1)
    acc = collections.defaultdict(list)
    for r in rows:
        name = r[h["Kernel Name"]]
        us = float(r[h["gpu__time_duration.sum"]])
        if us < 9.0:
            continue
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        if re.search(r"propagate_kernel<\w+, \w+, (\(bool\))?0", name):
            # step launches; a gathering one reads ancestors + state (12 B per particle), a plain one state + log weights (16 B)
            acc["_prop"].append((rd, wr, us))
            continue
        for tag in ("weights_kernel", "partition_kernel", "search_sorted_kernel"):
            if tag in name:
                acc[tag.replace("_kernel", "").replace("_sorted", "")].append((rd + wr, us))
    res = {"source": "%s (ncu --set full --clock-control none, `%s`): dram__bytes_read.sum + dram__bytes_write.sum per full-size launch" % (src.split("/")[-1], cmd)}
    prop = acc.get("_prop", [])
    if prop:
        top = max(p[0] for p in prop)
        plain = [p for p in prop if p[0] >= 0.88 * top]
        gath = [p for p in prop if p[0] < 0.88 * top]
        if plain:
            res["propagate_plain_bytes"] = sum(p[0] + p[1] for p in plain) / len(plain)
            res["propagate_plain_us"] = sum(p[2] for p in plain) / len(plain)
        if gath:
            res["propagate_gather_bytes"] = sum(p[0] + p[1] for p in gath) / len(gath)
            res["propagate_gather_us"] = sum(p[2] for p in gath) / len(gath)
    for k in ("weights", "partition", "search"):
        if acc.get(k):
            top_us = max(v[1] for v in acc[k])
            full_size = [v for v in acc[k] if v[1] >= 0.5 * top_us]      # early-exit launches of non-resampling steps drop out
            res[k + "_bytes"] = sum(v[0] for v in full_size) / len(full_size)
            res[k + "_us"] = sum(v[1] for v in full_size) / len(full_size)
            res[k + "_launches"] = len(full_size)
    res["note"] = ("N=2^24 fp64. Algorithmic bytes per launch: propagate plain 536.9 MB, gathering 469.8 MB (traffic is lower: the tail of the "
                   "written log-weight column is still in the 126 MB L2 when the kernel ends), weights 268.4 MB, search 201.3 MB")
    json.dump(res, open(dst, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")

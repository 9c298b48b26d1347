"""Instruction / stall-sample profile per CUDA source line of one kernel from an ncu report (needs -lineinfo + --import-source on).
  python scripts/line_profile.py report.ncu-rep kernel_regex [top_n] [instr|samples] [instance]
The source page lists, per profiled launch, one section per source file; the sections of one launch are joined here
(all files), the launch with the most instructions is shown unless `instance` picks one."""
import collections, csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
by = 1 if len(sys.argv) > 4 and sys.argv[4].startswith("s") else 0
pick = int(sys.argv[5]) if len(sys.argv) > 5 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
cur, h, first_file = None, {}, None
per_inst, srcs, names = [], {}, []
fn = ""
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        if first_file is None: first_file = r[1]
        if r[1] == first_file: per_inst.append(collections.OrderedDict()); names.append("")
        continue
    if r[0] in ("Kernel Name", "Function Name"):
        if per_inst: names[-1] = r[1]
        continue
    if r[0] == "Line No":
        h = {}
        for i, k in enumerate(r): h.setdefault(k, i)
        continue
    if not per_inst or len(r) < 8: continue
    try: ln = int(r[0])
    except ValueError: continue
    def num(x):
        try: return int(x)
        except ValueError: return 0
    n, sm = num(r[h["Instructions Executed"]]), num(r[h["# Samples"]])
    agg = per_inst[-1]
    k = (cur, ln)
    if k not in agg: agg[k] = [0, 0]; srcs[k] = r[1].strip()[:100]
    agg[k][0] += n; agg[k][1] += sm
idx = pick if pick is not None else max(range(len(per_inst)), key=lambda i: sum(v[0] for v in per_inst[i].values()))
agg = per_inst[idx]
tot, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print("launch %d of %d: %s" % (idx, len(per_inst), names[idx][:90]))
print("warp instructions", tot, "samples", ts)
files = collections.OrderedDict()
for (f, _), v in agg.items():
    a = files.setdefault(f, [0, 0]); a[0] += v[0]; a[1] += v[1]
for f, v in sorted(files.items(), key=lambda kv: -kv[1][0]):
    print("  %-32s %5.1f%% instr %5.1f%% samples" % (f, 100 * v[0] / max(tot, 1), 100 * v[1] / max(ts, 1)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
    print("%5.1f%% instr %5.1f%% samples  %s:%d  %s" % (100 * v[0] / max(tot, 1), 100 * v[1] / max(ts, 1), k[0], k[1], srcs[k]))

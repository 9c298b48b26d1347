"""Instruction / stall-sample profile per CUDA source line of one kernel from an ncu report (needs -lineinfo + --import-source on).
  python scripts/line_profile.py report.ncu-rep kernel_regex [top_n]"""
import collections, csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
cur, inst, h = None, 0, {}
agg, srcs = collections.OrderedDict(), {}
best = None
per_inst = []
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] in ("Kernel Name", "Function Name"):
        inst += 1; agg = collections.OrderedDict(); per_inst.append(agg); continue
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
    k = (cur, ln)
    if k not in agg: agg[k] = [0, 0]; srcs[k] = r[1].strip()[:100]
    agg[k][0] += n; agg[k][1] += sm
agg = max(per_inst, key=lambda a: sum(v[0] for v in a.values()))
tot, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print("warp instructions", tot, "samples", ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% instr %5.1f%% samples  %s:%d  %s" % (100 * v[0] / tot, 100 * v[1] / max(ts, 1), k[0], k[1], srcs[k]))

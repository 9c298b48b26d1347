"""Opcode / hot-spot profile of one kernel from an ncu report (source page, SASS).
  python scripts/sass_profile.py report.ncu-rep kernel_regex [units_per_launch]"""
import collections, csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
insts, cur, names = [], None, []
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name": cur = []; insts.append(cur); names.append(r[1]); continue
    if r and r[0] == "Address": hdr = r; continue
    if cur is not None and len(r) > 6: cur.append(r)
h = {k: i for i, k in enumerate(hdr)}
tots = [sum(int(r[h["Instructions Executed"]]) for r in o) for o in insts]
for k, (n, t) in enumerate(zip(names, tots)): print(k, t, n[:90])
sel = int(sys.argv[4]) if len(sys.argv) > 4 else max(range(len(insts)), key=lambda k: tots[k])
o, tot = insts[sel], tots[sel]
print("instance", sel, "warp instructions", tot, ("= %.1f thread-instr per unit" % (tot * 32 / units)) if units else "")
ops, samp = collections.Counter(), collections.Counter()
for r in o:
    n = int(r[h["Instructions Executed"]])
    w = [x for x in r[h["Source"]].strip().split() if not x.startswith("@")]
    op = w[0].split(".")[0]
    ops[op] += n; samp[op] += int(r[h["# Samples"]])
ts = sum(samp.values())
for op, n in ops.most_common(22):
    print(op.ljust(8), "%5.1f%% instr" % (100 * n / tot), ("%7.2f/unit" % (n * 32 / units)) if units else "", "%5.1f%% samples" % (100 * samp[op] / max(ts, 1)))

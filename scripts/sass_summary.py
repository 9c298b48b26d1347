"""SASS evidence from the built library (no GPU needed): per kernel, the instruction count and the mnemonics that show
what the code uses -- UBLKCP (cp.async.bulk = TMA bulk copy), SYNCS (mbarrier), ACQBULK / griddepcontrol (programmatic
dependent launch), DFMA/DMUL/DADD (fp64), IMAD.WIDE (Philox), LDS/STS, ATOM, SHFL, and that no tensor-core instruction
(HMMA / UTC*MMA) is present (nothing on this path is a contraction).
  python scripts/sass_summary.py > profiles/<tag>_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "gen_b200", "libgensmc.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
cur, funcs = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = funcs.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        cur[m.group(1).split(".")[0]] += 1
        cur["_total"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
keys = ["UBLKCP", "SYNCS", "ACQBULK", "DFMA", "DMUL", "DADD", "IMAD", "LDS", "STS", "SHFL", "ATOM", "ATOMS", "ATOMG", "RED", "HMMA", "UTCHMMA", "UTCQMMA"]
print("# SASS summary of gen_b200/libgensmc.so (cuobjdump -sass); architectures in the fatbin: %s\n" % ", ".join(archs))
print("| kernel | instructions | " + " | ".join(keys) + " |")
print("|---|---:|" + "---:|" * len(keys))
for (mangled, c), name in zip(funcs.items(), demangle):
    short = re.sub(r"\(.*$", "", name).replace("void ", "")
    if c["_total"] < 50:
        continue
    print("| `%s` | %d | %s |" % (short[:90], c["_total"], " | ".join(str(c[k]) if c[k] else "" for k in keys)))
tc = sum(c["HMMA"] + c["UTCHMMA"] + c["UTCQMMA"] for c in funcs.values())
print("\ntensor-core instructions in the library: %d (nothing on this path is a dense contraction)" % tc)
print("UBLKCP = cp.async.bulk (TMA bulk copy of the CDF window), SYNCS = mbarrier, ACQBULK = griddepcontrol (programmatic dependent launch)")

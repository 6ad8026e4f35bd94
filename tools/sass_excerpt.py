"""SASS evidence for profiles/: per-kernel counts of the instructions that prove TMA / mbarrier / cp.async use
(UTMALDG, SYNCS, LDGSTS, UTMASTG, MEMBAR, fp64 DFMA) in the in-tree libdppb200.so, plus the first lines around
each UTMALDG of the fused CG kernel.  Runs on the build host (cuobjdump only, no GPU)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "perphil_b200/libdppb200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, counts, ctx = None, collections.OrderedDict(), {}
PAT = ["UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "MEMBAR", "DFMA", "DADD", "DMUL", "LDS", "STG", "LDG", "ATOM", "RED", "ERRBAR", "CCTL", "ACQBULK", "UBLKCP"]
lines = out.splitlines()
for i, ln in enumerate(lines):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"dpp::\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern) or m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m:
        op = m.group(1)
        base = op.split(".")[0]
        if base in PAT:
            counts[kern][base] += 1
        if op.startswith("UTMALDG") and "k_cg_fused_apply<2, 1>" in kern and len(ctx.setdefault(kern, [])) < 6:
            ctx[kern].append(ln.strip())
print(f"# SASS excerpt of {so}\n")
print(f"cubin architectures: {', '.join(arch)}\n")
print("| kernel | " + " | ".join(PAT) + " |")
print("|---|" + "---:|" * len(PAT))
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    if any(c[p] for p in ("UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "MEMBAR", "DFMA")):
        print(f"| `{k}` | " + " | ".join(str(c[p]) for p in PAT) + " |")
print(f"| **all {len(counts)} kernels** | " + " | ".join(str(tot[p]) for p in PAT) + " |")
for k, ls in ctx.items():
    print(f"\n`{k}`: TMA tensor loads as emitted\n\n```")
    for l in ls:
        print(l)
    print("```")

"""Top stall locations of one kernel from an ncu report (source page; compile with -lineinfo).
usage: python tools/ncu_hotspots.py <report.ncu-rep> <kernel regex> [top N]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the report may hold several launches: keep the first block
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break
        hdr = r
        continue
    if hdr is not None and len(r) == len(hdr):
        data.append(r)
si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in data)
print("total samples", tot, "instructions", len(data), "warp-instr executed", sum(int(r[ie]) for r in data))
agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stall_cols}
print("by reason:", ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][si]))[:top_n]:
    st = sorted([(int(r[i]), hdr[i][6:]) for i in stall_cols], reverse=True)[:2]
    print(f"{idx:5d} {100*int(r[si])/tot:5.1f}% {r[ie]:>8} {r[src][:72]:72s} {st}")

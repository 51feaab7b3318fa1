"""Top stall sites of one kernel in an ncu report (SASS view): python scripts/ncu_hot.py <rep> <kernel regex> [n]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(H) and r[0] != "Address"]

def I(v):
    try:
        return int(v)
    except ValueError:
        return 0
si, ii = H.index("# Samples"), H.index("Instructions Executed")
stalls = [c for c in H if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(I(r[si]) for r in data)
print("kernel:", rows[0][1][:80], "total samples", tot, "instructions", sum(I(r[ii]) for r in data))
agg = {c: sum(I(r[H.index(c)]) for r in data) for c in stalls}
print("stall mix:", ", ".join(f"{k[6:]} {100*v/max(tot,1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -I(r[si]))[:n]:
    top = sorted(((c, I(r[H.index(c)])) for c in stalls), key=lambda kv: -kv[1])[:2]
    print(f"{100*I(r[si])/max(tot,1):5.1f}%  x{r[ii]:>7s}  {r[1].strip()[:70]:70s} {top[0][0][6:]}:{top[0][1]} {top[1][0][6:]}:{top[1][1]}")

"""Top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv [--kernel-name ...]` output."""
import csv, sys
path = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
def I(r, k):
    try: return int(r[ix[k]] or 0)
    except ValueError: return 0
tot = sum(I(r, "# Samples") for r in data)
print(rows[0][1][:100]); print("total samples", tot, "instructions", sum(I(r, "Instructions Executed") for r in data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(I(r, s) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {s}: {v} ({100*v/max(tot,1):.1f}%)")
for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:n]:
    st = {s: I(r, s) for s in stalls}
    m = max(st, key=st.get)
    print(f'{I(r,"# Samples"):6d} {100*I(r,"# Samples")/max(tot,1):5.1f}% exec={r[ix["Instructions Executed"]]:>9} {m:16s} {r[ix["Source"]].strip()[:100]}')

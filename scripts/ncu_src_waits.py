"""Where the warps of a kernel wait: from an `ncu --page source --csv` export, the stall samples of every mbarrier
try_wait loop (mapped to the barrier's shared-memory offset) and the rest grouped by opcode family.
usage: python scripts/ncu_src_waits.py src.csv [bars_offset_hex]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr)]
def samples(r):
    try: return int(r[idx['# Samples']])
    except ValueError: return 0
tot = sum(samples(r) for r in body)
waits = collections.Counter(); other = collections.Counter()
last_sync = None; since = 99
for r in body:
    src = r[idx['Source']].strip()
    m = re.search(r'SYNCS\.PHASECHK\.TRANS64\.TRYWAIT P\d, \[(.*?)\]', src)
    if m:
        last_sync = m.group(1).split('+')[-1]; since = 0
    s = samples(r)
    if since <= 3 and last_sync is not None and ('BRA' in src or 'SYNCS' in src or 'ISETP' in src or 'NOP' in src):
        waits[last_sync] += s
    else:
        op = re.sub(r'^@!?U?P\d+\s+', '', src).split(' ')[0].split('.')[0]
        other[op] += s
    since += 1
print(f"total samples {tot}")
base = int(sys.argv[2], 16) if len(sys.argv) > 2 else None
for k, v in waits.most_common(12):
    tag = ''
    if base is not None:
        try: tag = f" (bars[{(int(k, 16) - base) // 8}])"
        except ValueError: pass
    print(f"  wait on mbarrier @{k}{tag}: {v} ({100.0 * v / tot:.1f} %)")
print("other, by opcode:")
for k, v in other.most_common(16):
    print(f"  {k:10s} {v} ({100.0 * v / tot:.1f} %)")

"""Top stall sites of a kernel from an ncu report's source page (SASS view).
    python tools/ncu_hot.py report.ncu-rep [topN]
"""
import csv, io, subprocess, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# first kernel only
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == 'Address':
        break
    body.append(r)
tot = sum(int(r[idx['# Samples']]) for r in body)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print("kernel:", rows[0][1][:120], " total samples", tot)
agg = {c: sum(int(r[idx[c]] or 0) for r in body) for c in stall_cols}
print("by reason:", ", ".join("%s=%.1f%%" % (c[6:], 100.0 * v / tot) for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
ranked = sorted(range(len(body)), key=lambda i: -int(body[i][idx['# Samples']]))[:top]
for i in sorted(ranked):
    r = body[i]
    reasons = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%%  %-60s %s" % (i, 100.0 * int(r[idx['# Samples']]) / tot, r[idx['Source']].strip()[:60],
                                   " ".join("%s:%d" % (n, v) for v, n in reasons if v)))

"""Hot SASS regions of a kernel from `ncu --page source --csv --print-source sass`:
   python profiles/sass_hot.py <sass.csv> <kernel index> [top N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
k = int(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lo = hdr[k] + 1; hi = hdr[k + 1] - 3 if k + 1 < len(hdr) else len(rows)
H = rows[hdr[k]]
ci, si, ei = H.index('Instructions Executed'), H.index('# Samples'), H.index('Source')
body = [r for r in rows[lo:hi] if len(r) > ci and r[ci].strip().isdigit()]
tot = sum(int(r[ci]) for r in body); tots = sum(int(r[si] or 0) for r in body)
print("instructions", len(body), "executed", tot, "samples", tots)
# contiguous regions of similar execution count
reg = []
for idx, r in enumerate(body):
    e = int(r[ci])
    if reg and 0.5 < (e + 1) / (reg[-1]['e'] + 1) < 2.0:
        reg[-1]['n'] += 1; reg[-1]['sum'] += e; reg[-1]['s'] += int(r[si] or 0); reg[-1]['ops'].append(r[ei].split()[0] if r[ei].split() else '')
    else:
        reg.append(dict(start=idx, n=1, e=e, sum=e, s=int(r[si] or 0), ops=[r[ei].split()[0] if r[ei].split() else '']))
for g in sorted(reg, key=lambda g: -g['s'])[:top]:
    import collections
    c = collections.Counter(o.split('.')[0] for o in g['ops']).most_common(6)
    print("@%5d n=%4d exec/instr=%10d  instr%%=%5.1f samples%%=%5.1f  %s" % (g['start'], g['n'], g['e'], 100.0 * g['sum'] / tot, 100.0 * g['s'] / max(tots, 1), c))

"""Top CUDA source lines of an `ncu --page source --csv --print-source cuda,sass` dump: python tools/srcprof.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows[:8]) if 'Line No' in r)
hdr = rows[hi]
ln, src, ie, smp = hdr.index('Line No'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
def I(x):
    try: return int(x)
    except Exception: return 0
items = [(I(r[smp]), I(r[ie]), r[ln], r[src].strip()[:105]) for r in rows[hi + 1:] if len(r) >= len(hdr) and r[ln].strip() not in ('', '-')]
ts = sum(i[0] for i in items) or 1; ti = sum(i[1] for i in items) or 1
print('samples', ts, 'warp-instr', ti)
for s_, n, l, t in sorted(items, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print('smp %5.1f%% ins %5.1f%%  L%-4s %s' % (100 * s_ / ts, 100 * n / ti, l, t))

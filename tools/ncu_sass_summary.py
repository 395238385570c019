"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` output: stall mix and the
hottest SASS instructions of the first kernel instance.  Usage: ncu_sass_summary.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
d = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
def f(r, h):
    try:
        return float(r[ix[h]] or 0)
    except ValueError:
        return 0.0
tot_inst = sum(f(r, 'Instructions Executed') for r in d)
tot_samp = sum(f(r, '# Samples') for r in d)
print("kernel:", rows[hi - 1][1][:120] if hi else "")
print("sass lines", len(d), "warp instructions", tot_inst, "samples", tot_samp)
for h in hdr:
    if h.startswith('stall_') and 'Not Issued' not in h:
        v = sum(f(r, h) for r in d)
        if v > 0.01 * tot_samp:
            print(f"  {h:28s} {v:10.0f} {v / max(tot_samp, 1) * 100:5.1f}%")
print("top by samples: samples / warp-inst / avg threads / addr / sass")
for r in sorted(d, key=lambda r: -f(r, '# Samples'))[:top]:
    print(r[ix['# Samples']].rjust(7), r[ix['Instructions Executed']].rjust(9), r[ix['Avg. Threads Executed']].rjust(6), r[0][-5:], r[1][:100])

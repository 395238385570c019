"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line:
samples (barrier / long-scoreboard share), warp instructions, average active threads.  Usage: ncu_lines.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) != len(hdr) or r[0] == '' or r[2] != '-': continue
    def f(h):
        try: return float(r[ix[h]] or 0)
        except ValueError: return 0.0
    agg[(cur, int(r[0]))] = (f('# Samples'), f('Instructions Executed'), f('Thread Instructions Executed'), f('stall_barrier'), f('stall_long_sb'), r[1].strip()[:80])
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values()); tott = sum(v[2] for v in agg.values())
print(f'total samples {tot:.0f}  warp instructions {toti:.0f}  avg active threads {tott / max(toti, 1):.1f}')
print('--- by samples')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]:16s}:{k[1]:4d} samp {v[0]:6.0f} ({100 * v[0] / tot:4.1f}%; barrier {v[3]:5.0f} long_sb {v[4]:5.0f}) inst {v[1]:9.0f} thr {v[2] / max(v[1], 1):4.1f} | {v[5]}")
print('--- by warp instructions')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]:16s}:{k[1]:4d} inst {v[1]:9.0f} ({100 * v[1] / toti:4.1f}%) thr {v[2] / max(v[1], 1):4.1f} samp {v[0]:6.0f} | {v[5]}")

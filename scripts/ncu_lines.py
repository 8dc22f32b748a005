"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line (top N by stall samples)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur = None; ix = None; out = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        ix = {}
        for i, h in enumerate(r): ix.setdefault(h, i)
        continue
    if ix and r[0].isdigit():
        g = lambda k: int(r[ix[k]] or 0)
        out.append((g('# Samples'), cur, int(r[0]), r[1].strip()[:90], g('Instructions Executed'), g('stall_long_sb'), g('stall_no_inst'), g('stall_wait'), g('stall_short_sb')))
out.sort(reverse=True)
print('samples file line | instr long_sb no_inst wait short_sb')
for o in out[:top]: print(o[0], o[1][-16:], o[2], '|', o[4], o[5], o[6], o[7], o[8], '|', o[3])

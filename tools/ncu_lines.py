"""Rank CUDA source lines of an ncu report by executed instructions / stall samples.
usage: python tools/ncu_lines.py <cuda,sass csv> [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur = None; hdr = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; hdr = None; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == '-':          # a CUDA source line summary row
        ii = hdr.index('Instructions Executed'); si = hdr.index('# Samples')
        try:
            out.append((int(r[ii]), int(r[si]), cur, r[0], r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(o[0] for o in out); tots = sum(o[1] for o in out)
print('total warp-inst', tot, 'samples', tots)
key = 1 if len(sys.argv) > 3 and sys.argv[3] == 'samples' else 0
out.sort(key=lambda o: -o[key])
for n, sm, f, l, src in out[:top]:
    print(f'{n:9d} {100*n/max(tot,1):5.1f}%  samp {sm:6d} {100*sm/max(tots,1):5.1f}%  {f}:{l}  {src}')

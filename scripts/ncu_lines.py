"""Aggregate an ncu source page (cuda,sass) by CUDA source line: samples, instructions."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
fname = None; hdr = None; data = []
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; si = r.index('# Samples'); ii = r.index('Instructions Executed'); continue
    if hdr is None or len(r) <= ii: continue
    if r[0] in ('-', ''): continue   # SASS rows under a source line
    try: data.append((int(r[si]), int(r[ii]), fname, int(r[0]), r[1]))
    except ValueError: pass
tot = sum(d[0] for d in data) or 1; toti = sum(d[1] for d in data) or 1
print('total samples', tot, 'warp-instructions', toti)
for s, i, f, ln, src in sorted(data, reverse=True)[:topn]:
    print(f'{100*s/tot:5.1f}% smp {100*i/toti:5.1f}% inst  {f}:{ln}  {src.strip()[:100]}')

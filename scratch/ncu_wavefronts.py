"""Per CUDA line: shared-memory instructions, wavefronts and wavefronts per instruction of an ncu report (source page)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; iters = float(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 1000 / 3
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
hdr = None; f = None; cur = None
agg = collections.OrderedDict()
def I(x):
    try: return int(x)
    except Exception: return 0
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'File Path': f = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; iw = hdr.index('L1 Wavefronts Shared'); ii = hdr.index('Instructions Executed'); iid = hdr.index('L1 Wavefronts Shared Ideal'); continue
    if hdr is None or len(r) <= iw: continue
    if r[0].isdigit():
        cur = (f, int(r[0]), r[1].strip()[:80])
    elif r[0] == '' and cur is not None and len(r) > 3:
        t = r[3].split()
        if not t: continue
        op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
        if op.startswith(('LDS', 'STS')):
            a = agg.setdefault(cur, {})
            k = op.split('.')[0] + ('.128' if '128' in op else ('.64' if '64' in op else ''))
            e = a.setdefault(k, [0, 0, 0])
            e[0] += I(r[ii]); e[1] += I(r[iw]); e[2] += I(r[iid])
tot = 0
for (f, ln, text), a in agg.items():
    for k, (n, w, wi) in a.items():
        if n == 0: continue
        tot += w
        print(f"{f}:{ln:<4d} {k:8s} n={n/iters:6.1f}/it  wavefronts={w/iters:7.1f}/it  per instr {w/n:4.2f} (ideal {wi/n:4.2f})  {text}")
print("total wavefronts per warp-iteration", tot / iters)

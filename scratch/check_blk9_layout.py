"""Independent wavefront count for the constants in c3_b200/csrc/pwc_blk9.cuh (true 16-byte addresses)."""
import re, sys
src = open('/root/repo/c3_b200/csrc/pwc_blk9.cuh').read()
def tab(name):
    m = re.search(name + r"\[\d+\] = \{([^}]*)\}", src); return [int(x) for x in m.group(1).replace('\n', ' ').split(',')]
def wf(addrs):
    tot = 0
    for q in range(4):
        slots = {}
        for x in addrs[q*8:(q+1)*8]:
            if x is None: continue
            slots.setdefault(x % 8, set()).add(x)
        tot += max((len(v) for v in slots.values()), default=0)
    return tot
S = 9; BUF = 81
for nosel in (0, 1):
    pre = 'kB9n' if nosel else 'kB9'
    slot, perm, shadow, kord = tab(pre + 'Slot'), tab(pre + 'Perm'), tab(pre + 'Shadow'), tab(pre + 'Kord')
    g1 = int(re.search(r"G1 = NOSEL \? (\d+) : (\d+)", src).group(1 if nosel else 2)); g2 = int(re.search(r"G2 = NOSEL \? (\d+) : (\d+)", src).group(1 if nosel else 2))
    goff = [0, g1, g2]
    assert sorted(slot) == list(range(9)) and sorted(perm) == list(range(27))
    lanes = []
    for lane in range(32):
        s = lane if lane < 27 else shadow[lane-27]
        g = perm[s] // 9; li = perm[s] % 9; bi, bj = li // 3, li % 3
        yd = None
        if bi != bj: kx1, ky1, k2 = bi, bj, 3 - bi - bj
        else:
            kx1 = ky1 = (bi + 1 + kord[s]) % 3; k2 = (bi + 2 - kord[s]) % 3
            if nosel: yd = slot[ky1*3+bj]; ky1 = bi
        lanes.append(dict(g=g, on=lane < 27, own=slot[li], x1=slot[bi*3+kx1], y1=slot[ky1*3+bj], x2=slot[bi*3+k2], y2=slot[k2*3+bj], yd=yd))
    print("NOSEL" if nosel else "SEL", "group bases mod 8:", [x % 8 for x in goff])
    for name in ['x1', 'y1', 'x2', 'y2'] + (['yd'] if nosel else []):
        print(' ', name, [wf([goff[l['g']] + 2*BUF + e*S + l[name] if l[name] is not None else None for l in lanes]) for e in range(9)])
    print('  store', [wf([goff[l['g']] + 1*BUF + e*S + l['own'] if l['on'] else None for l in lanes]) for e in range(9)])
    print('  ownld', [wf([goff[l['g']] + 3*BUF + e*S + l['own'] for l in lanes]) for e in range(9)])
    print('  model', [wf([e*S + l['own'] for l in lanes]) for e in range(9)])

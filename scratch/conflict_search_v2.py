"""Bank-slot conflict model for the own-block product (pwc_blk v2): 16-byte slots, 8 per 128-byte wavefront,
quarter-warp = 8 consecutive lanes; identical addresses merge.  Searches LD, group offsets, lane rotations."""
import itertools, sys
import numpy as np

def wavefronts(addrs):  # addrs: list of 32 slot addresses (16-byte units) or None
    tot = 0
    for q in range(4):
        a = set(x for x in addrs[q*8:(q+1)*8] if x is not None)
        slots = {}
        for x in a:
            slots.setdefault(x % 8, set()).add(x)
        tot += max((len(v) for v in slots.values()), default=0)
    return tot

def lane_blocks(rots, perm=None):
    """lane -> (g, bi, bj)"""
    out = []
    for lane in range(32):
        g_raw = lane // 9
        on = g_raw < 3
        g = g_raw if on else 2
        li = ((lane - g_raw*9) if on else 8)
        li = (li + rots[g]) % 9
        out.append((g, li // 3, li % 3, on))
    return out

def evaluate(LD, offs, rots, diag_swap=0, verbose=False):
    lanes = lane_blocks(rots)
    total = 0; n = 0
    worst = 0
    def instr(fn):
        nonlocal total, n, worst
        addrs = []
        for (g, bi, bj, on) in lanes:
            r, c = fn(bi, bj)
            addrs.append(offs[g] + r*LD + c)
        w = wavefronts(addrs)
        total += w; n += 1; worst = max(worst, w)
    def k1(bi, bj):
        if bi != bj: return None
        return (bi + 1 + diag_swap) % 3 if diag_swap == 0 else (bi + 2) % 3
    for a in range(3):
        for kk in range(3):
            # LX1
            instr(lambda bi, bj: (3*bi+a, 3*bi+kk) if bi != bj else (3*bi+a, 3*((bi+1+diag_swap)%3)+kk))
            # LX2
            instr(lambda bi, bj: (3*bi+a, 3*(3-bi-bj)+kk) if bi != bj else (3*bi+a, 3*((bi+2-diag_swap)%3)+kk))
    for kk in range(3):
        for b in range(3):
            instr(lambda bi, bj: (3*bj+kk, 3*bj+b) if bi != bj else (3*((bi+1+diag_swap)%3)+kk, 3*bj+b))
            instr(lambda bi, bj: (3*(3-bi-bj)+kk, 3*bj+b) if bi != bj else (3*((bi+2-diag_swap)%3)+kk, 3*bj+b))
    ld_total, ld_n = total, n
    # stores of own block (only on-lanes; shadow lanes don't store)
    st_total = 0
    for a in range(3):
        for b in range(3):
            addrs = []
            for (g, bi, bj, on) in lanes:
                addrs.append(offs[g] + (3*bi+a)*LD + 3*bj+b if on else None)
            st_total += wavefronts(addrs)
    # legacy full loads (x row-block / y col-block) for the accumulate product
    leg = 0
    for a in range(3):
        for k in range(9):
            leg += wavefronts([offs[g] + (3*bi+a)*LD + k for (g, bi, bj, on) in lanes])
    for k in range(9):
        for b in range(3):
            leg += wavefronts([offs[g] + k*LD + 3*bj+b for (g, bi, bj, on) in lanes])
    return ld_total / ld_n, st_total / 9, leg / 54, worst

if __name__ == "__main__":
    print("current layout:", evaluate(11, (0, 405, 803), (0, 2, 0)))
    best = []
    for LD in range(9, 17):
        for o1 in range(8):
            for o2 in range(8):
                for r1 in range(9):
                    for r2 in range(9):
                        for ds in (0, 1):
                            l, s, leg, w = evaluate(LD, (0, o1, o2), (0, r1, r2), ds)
                            best.append((l + 0.25*s + 0.1*leg, l, s, leg, w, LD, o1, o2, r1, r2, ds))
    best.sort()
    for b in best[:15]: print(b)

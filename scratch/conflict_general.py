"""Like conflict_blockmajor.py (NOSEL pattern set) but with a free lane -> (group, block) assignment over the whole warp and
all shared-memory access patterns of pwc_blk9_t18_kernel weighted by their count per slice."""
import random, sys, math
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)

def wf(keys):
    tot = 0
    for q in range(4):
        slots = {}
        for k in keys[q*8:(q+1)*8]:
            if k is None: continue
            slots.setdefault(k[1], set()).add(k[0])
        tot += max((len(v) for v in slots.values()), default=0)
    return tot

W = dict(x1=54, y1=54, x2=54, y2=54, yd=54, store=117, ownld=36, model=27)

def cost(st, detail=False):
    res, offs, assign, shadow, kord = st
    lanes = []
    for lane in range(32):
        src = lane if lane < 27 else shadow[lane - 27]
        v = assign[src]; g = v // 9; li = v % 9
        lanes.append((g, li // 3, li % 3, lane < 27, src))
    pats = {k: [] for k in W}
    for (g, bi, bj, on, src) in lanes:
        def key(b, merge_groups=False):
            blk = b[0]*3 + b[1]
            if merge_groups: return ((0, blk), res[blk])
            return ((g, blk), (res[blk] + offs[g]) & 7)
        if bi != bj:
            k2 = 3 - bi - bj
            pats['x1'].append(key((bi, bi))); pats['y1'].append(key((bj, bj))); pats['yd'].append(None)
        else:
            k1 = (bi + 1 + kord[src]) % 3; k2 = (bi + 2 - kord[src]) % 3
            pats['x1'].append(key((bi, k1))); pats['y1'].append(key((bi, bi))); pats['yd'].append(key((k1, bi)))
        pats['x2'].append(key((bi, k2))); pats['y2'].append(key((k2, bj)))
        pats['store'].append(key((bi, bj)) if on else None)
        pats['ownld'].append(key((bi, bj)))
        pats['model'].append(key((bi, bj), True))
    d = {k: wf(v) for k, v in pats.items()}
    if detail: return d
    return sum(W[k] * d[k] for k in W)

def rand_state():
    return [[random.randrange(8) for _ in range(9)], [0, random.randrange(8), random.randrange(8)],
            random.sample(range(27), 27), [random.randrange(27) for _ in range(5)], [random.randrange(2) for _ in range(27)]]

def mutate(st):
    res, offs, assign, shadow, kord = st
    st = [list(res), list(offs), list(assign), list(shadow), list(kord)]
    m = random.random()
    if m < 0.2: st[0][random.randrange(9)] = random.randrange(8)
    elif m < 0.28: st[1][random.randrange(1, 3)] = random.randrange(8)
    elif m < 0.8:
        i, j = random.sample(range(27), 2); st[2][i], st[2][j] = st[2][j], st[2][i]
    elif m < 0.9: st[3][random.randrange(5)] = random.randrange(27)
    else: st[4][random.randrange(27)] ^= 1
    return st

best = None
for restart in range(3):
    cur = rand_state(); cc = cost(cur)
    T = 80.0
    for it in range(150000):
        nx = mutate(cur); nc = cost(nx)
        if nc <= cc or random.random() < math.exp((cc - nc) / T):
            cur, cc = nx, nc
            if best is None or cc < best[0]:
                best = (cc, cur)
        T = max(1.0, T * 0.99995)
    print("restart", restart, "best", best[0], cost(best[1], True), flush=True)
print(best)

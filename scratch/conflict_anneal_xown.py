"""Layout search for the X-own product: lane (bi,bj) takes X(bi,bj) from registers (k-block bj) and loads
Y(bj,bj), then for t = 1, 2 with kb = (bj + t) % 3 loads X(bi,kb) and Y(kb,bj) (per-lane kk rotation allowed)."""
import random, sys, math
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
FULL_PERM = len(sys.argv) > 2 and sys.argv[2] == "perm"

def wf(addrs):
    tot = 0
    for q in range(4):
        slots = {}
        for x in addrs[q*8:(q+1)*8]:
            if x is None: continue
            slots.setdefault(x & 7, set()).add(x)
        tot += max((len(v) for v in slots.values()), default=0)
    return tot

def cost(st, detail=False):
    LD, offs, perms, shadow, srot = st
    lanes = []
    for lane in range(32):
        src = lane if lane < 27 else shadow[lane - 27]
        g = src // 9
        li = perms[g][src % 9]
        lanes.append((g, li // 3, li % 3, lane < 27, src))
    ld = 0; n = 0
    for kk in range(3):
        for b in range(3):   # t = 0 : Y(bj,bj)[kk][b]
            ld += wf([(g << 12) + offs[g] + (3*bj+kk)*LD + 3*bj+b for (g, bi, bj, on, src) in lanes]); n += 1
    for t in (1, 2):
        for a in range(3):
            for kk in range(3):
                X, Y = [], []
                for (g, bi, bj, on, src) in lanes:
                    kb = (bj + t) % 3; r = (kk + srot[src][t-1]) % 3
                    X.append((g << 12) + offs[g] + (3*bi+a)*LD + 3*kb + r)                       # X(bi,kb)[a][r]
                    Y.append((g << 12) + offs[g] + (3*kb + (a + srot[src][t-1]) % 3)*LD + 3*bj + kk)   # Y(kb,bj)[r'][b]  (a,kk reused as r',b)
                ld += wf(X) + wf(Y); n += 2
    stt = 0
    for a in range(3):
        for b in range(3):
            stt += wf([(g << 12) + offs[g] + (3*bi+a)*LD + 3*bj+b if on else None for (g, bi, bj, on, src) in lanes])
    if detail: return ld / n, stt / 9
    return ld + stt

def rand_state():
    return [random.choice([9, 10, 11, 12, 13, 14, 15]), [0, random.randrange(8), random.randrange(8)],
            [random.sample(range(9), 9) if FULL_PERM else [(i + r) % 9 for i in range(9)] for r in [random.randrange(9) for _ in range(3)]],
            [26] * 5 if not FULL_PERM else [random.randrange(27) for _ in range(5)],
            [[0, 0] for _ in range(27)]]

def mutate(st):
    LD, offs, perms, shadow, srot = st
    st = [LD, list(offs), [list(p) for p in perms], list(shadow), [list(s) for s in srot]]
    m = random.random()
    if m < 0.05: st[0] = random.choice([9, 10, 11, 12, 13, 14, 15])
    elif m < 0.15: st[1][random.randrange(1, 3)] = random.randrange(8)
    elif m < 0.5:
        g = random.randrange(3)
        if FULL_PERM:
            p = st[2][g]; i, j = random.sample(range(9), 2); p[i], p[j] = p[j], p[i]
        else:
            r = random.randrange(9); st[2][g] = [(i + r) % 9 for i in range(9)]
    elif m < 0.6: st[3][random.randrange(5)] = random.randrange(27)
    else: st[4][random.randrange(27)][random.randrange(2)] = random.randrange(3)
    return st

best = None
for restart in range(3):
    cur = rand_state(); cc = cost(cur)
    T = 3.0
    for it in range(60000):
        nx = mutate(cur); nc = cost(nx)
        if nc <= cc or random.random() < math.exp((cc - nc) / T):
            cur, cc = nx, nc
            if best is None or cc < best[0]:
                best = (cc, cur)
        T = max(0.05, T * 0.9999)
    print("restart", restart, "best", best[0], cost(best[1], True), flush=True)
print(best)

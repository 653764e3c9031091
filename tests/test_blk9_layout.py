"""Shared-memory layout of the d = 9 own-block kernel (c3_b200/csrc/pwc_blk9.cuh): the lane -> (group, block) tables, block
slots and group bases in the header must give the wavefront counts DESIGN.md claims.  Bank model: 16-byte slots, 8 per
128-byte wavefront, a quarter-warp (8 consecutive lanes) per wavefront, identical addresses merge (ncu confirms it:
profiles/r01_prof_blk9_own_final.txt, 2.9 M conflict wavefronts of 2.35 G)."""
import os
import re

import pytest

HDR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "c3_b200", "csrc", "pwc_blk9.cuh")
S, BUF = 9, 81


def _tab(src, name):
    m = re.search(name + r"\[\d+\] = \{([^}]*)\}", src)
    return [int(x) for x in m.group(1).replace("\n", " ").split(",")]


def _wavefronts(addrs):
    tot = 0
    for q in range(4):
        slots = {}
        for x in addrs[q * 8:(q + 1) * 8]:
            if x is not None:
                slots.setdefault(x % 8, set()).add(x)
        tot += max((len(v) for v in slots.values()), default=0)
    return tot


@pytest.mark.parametrize("nosel", [True, False])
def test_blk9_tables_are_conflict_free(nosel):
    src = open(HDR).read()
    pre = "kB9n" if nosel else "kB9"
    slot, perm, shadow, kord = (_tab(src, pre + n) for n in ("Slot", "Perm", "Shadow", "Kord"))
    pick = 1 if nosel else 2
    g1 = int(re.search(r"G1 = NOSEL \? (\d+) : (\d+)", src).group(pick))
    g2 = int(re.search(r"G2 = NOSEL \? (\d+) : (\d+)", src).group(pick))
    we = int(re.search(r"WARP_ELEMS = (\d+);", src).group(1))
    nbuf = int(re.search(r"NBUF = (\d+);", src).group(1))
    assert sorted(slot) == list(range(9)) and sorted(perm) == list(range(27)) and all(0 <= x < 27 for x in shadow)
    assert g1 >= nbuf * BUF and g2 >= g1 + nbuf * BUF and we >= g2 + nbuf * BUF and we % 8 == 0   # buffers of the groups do not overlap
    goff = [0, g1, g2]
    lanes = []
    for lane in range(32):
        s = lane if lane < 27 else shadow[lane - 27]
        g, li = divmod(perm[s], 9)
        bi, bj = divmod(li, 3)
        yd = None
        if bi != bj:
            kx1, ky1, k2 = bi, bj, 3 - bi - bj
        else:
            kx1 = ky1 = (bi + 1 + kord[s]) % 3
            k2 = (bi + 2 - kord[s]) % 3
            if nosel:
                yd, ky1 = slot[ky1 * 3 + bj], bi
        lanes.append(dict(g=g, on=lane < 27, own=slot[li], x1=slot[bi * 3 + kx1], y1=slot[ky1 * 3 + bj],
                          x2=slot[bi * 3 + k2], y2=slot[k2 * 3 + bj], yd=yd))
    # every lane pair (group, block) is owned exactly once
    assert len({(l["g"], l["own"]) for l in lanes[:27]}) == 27

    def wf(name, buf=2, stores=False):
        return [_wavefronts([goff[l["g"]] + buf * BUF + e * S + l[name]
                             if l[name] is not None and (l["on"] or not stores) else None for l in lanes]) for e in range(9)]

    for name in ("x1", "y1", "x2", "y2"):          # the four operand loads of a product: the hardware minimum
        assert wf(name) == [4] * 9, (name, wf(name))
    assert wf("own", buf=1, stores=True) == [4] * 9                                      # publishing stores
    if nosel:
        assert wf("yd") == [3] * 9                                                       # diagonal-lane fetch (9 lanes)
        assert wf("own", buf=3) == [4] * 9                                               # own-block reloads (all 32 lanes)
        assert [_wavefronts([e * S + l["own"] for l in lanes]) for e in range(9)] == [4] * 9   # generators (shared buffer)

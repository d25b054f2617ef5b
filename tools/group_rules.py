"""CPU replay of target-group rules: number of groups (what the pair kernels' cost is proportional to), lane fill and candidates per
group (particles of the leaf cells whose box lies within 2h x 1.02 of the group's box, minimum image) -- to choose a grouping rule
before spending GPU time on it.  Calibration against the B200 (100^3 lattice): plain subtrees 400.3 candidates, Morton-run packing
inside 256-particle subtrees 572.7.

usage: python tools/group_rules.py [nx] [lattice|random] [leafscan]     (leafscan: candidates per group against max_leaf)"""
import sys
import numpy as np
from scipy.spatial import cKDTree
from group_fill import part1by2


def prefix_nodes(key, limit):
    """id of the maximal key-prefix node with <= limit particles holding each (sorted) particle"""
    out = np.zeros(key.size, dtype=np.int64)
    todo = np.arange(key.size)
    base = 0
    for L in range(1, 49):
        pre = key[todo] >> np.uint64(48 - L)
        u, inv, cnt = np.unique(pre, return_inverse=True, return_counts=True)
        done = cnt[inv] <= limit
        out[todo[done]] = base + inv[done]
        base += len(u)
        todo = todo[~done]
        if todo.size == 0:
            break
    return out


def runs(ids):
    s = np.flatnonzero(np.r_[True, ids[1:] != ids[:-1]])
    return s, np.diff(np.r_[s, ids.size])


def boxes(pos, starts):
    return np.minimum.reduceat(pos, starts, axis=0), np.maximum.reduceat(pos, starts, axis=0)


def candidates(glo, ghi, clo, chi, ccount, r):
    """particles in cells whose box is within r of each group's box (periodic unit box)"""
    cc, ch = 0.5 * (clo + chi), 0.5 * (chi - clo)
    gc, gh = 0.5 * (glo + ghi), 0.5 * (ghi - glo)
    tree = cKDTree(np.mod(cc, 1.0), boxsize=1.0)
    rad = r + np.linalg.norm(gh, axis=1) + np.linalg.norm(ch, axis=1).max()
    out = np.zeros(len(gc))
    for g in range(len(gc)):
        idx = np.array(tree.query_ball_point(np.mod(gc[g], 1.0), rad[g]))
        d = np.abs(cc[idx] - gc[g]); d = np.minimum(d, 1.0 - d)
        gap = np.maximum(d - ch[idx] - gh[g], 0.0)
        out[g] = ccount[idx[(gap ** 2).sum(1) <= r * r]].sum()
    return out


def merge_rule(gstart, gcount, glo, ghi, cap, r, grow):
    """greedy merge of Morton-consecutive groups while the sum fits a warp and the search volume of the union stays within
    `grow` x the larger of the two search volumes"""
    def vol(lo, hi):
        return np.prod(hi - lo + 2 * r)
    S, C, LO, HI = [gstart[0]], [gcount[0]], [glo[0]], [ghi[0]]
    for k in range(1, len(gstart)):
        lo, hi = np.minimum(LO[-1], glo[k]), np.maximum(HI[-1], ghi[k])
        if C[-1] + gcount[k] <= cap and vol(lo, hi) <= grow * max(vol(LO[-1], HI[-1]), vol(glo[k], ghi[k])):
            C[-1] += gcount[k]; LO[-1], HI[-1] = lo, hi
        else:
            S.append(gstart[k]); C.append(gcount[k]); LO.append(glo[k]); HI.append(ghi[k])
    return np.array(S), np.array(C), np.array(LO), np.array(HI)


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    kind = sys.argv[2] if len(sys.argv) > 2 else "lattice"
    if kind == "lattice":
        g = (np.arange(nx) + 0.5) / nx
        pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    else:
        pos = np.random.default_rng(1).random((nx ** 3, 3))
    q = np.minimum((pos * 65536).astype(np.int64), 65535)
    key = (part1by2(q[:, 0]) << np.uint64(2)) | (part1by2(q[:, 1]) << np.uint64(1)) | part1by2(q[:, 2])
    order = np.argsort(key, kind="stable")
    key, pos = key[order], pos[order]
    r = 2.0 * 1.2 / nx * 1.02
    cs, ccount = runs(prefix_nodes(key, 8))
    clo, chi = boxes(pos, cs)
    gs, gcount = runs(prefix_nodes(key, 32))
    glo, ghi = boxes(pos, gs)
    sample = np.random.default_rng(2).choice(len(gs), size=min(3000, len(gs)), replace=False)

    def report(name, S, C, LO, HI):
        smp = np.random.default_rng(2).choice(len(S), size=min(3000, len(S)), replace=False)
        cand = candidates(LO[smp], HI[smp], clo, chi, ccount, r)
        print(f"{name:34s} groups {len(S):7d}  fill {C.mean():5.2f}/32  candidates mean {cand.mean():6.1f}  p95 {np.percentile(cand, 95):6.1f}"
              f"  > 384: {100 * (cand > 384).mean():4.1f} %  > 512: {100 * (cand > 512).mean():4.1f} %  max {cand.max():.0f}")
    if len(sys.argv) > 3 and sys.argv[3] == "leafscan":
        smp = np.random.default_rng(2).choice(len(gs), size=min(2000, len(gs)), replace=False)
        for leaf in (2, 4, 8, 16):
            ls, lcount = runs(prefix_nodes(key, leaf))
            llo, lhi = boxes(pos, ls)
            cand = candidates(glo[smp], ghi[smp], llo, lhi, lcount, r)
            print(f"max_leaf {leaf:2d}: cells {len(ls):7d} (mean {lcount.mean():5.2f} particles)  candidates mean {cand.mean():6.1f}  p95 {np.percentile(cand, 95):6.1f}"
                  f"  cells per list {cand.mean() / lcount.mean():5.0f}  > 384: {100 * (cand > 384).mean():4.1f} %")
        return
    report("plain (subtrees <= 32)", gs, gcount, glo, ghi)
    for grow in (1.3, 1e9):
        report(f"merge neighbours, volume x{grow:g}", *merge_rule(gs, gcount, glo, ghi, 32, r, grow))


if __name__ == "__main__":
    sys.path.insert(0, __file__.rsplit("/", 1)[0])
    main()

"""Replay of the target-group rule (maximal binary-radix subtrees with <= 32 particles over 48-bit Morton keys, tree.cu k_groups) in numpy:
mean / count / percentiles of the group sizes for lattices and random positions (target-group fill)."""
import numpy as np, sys
def part1by2(x):
    x = x.astype(np.uint64) & np.uint64(0xFFFF)
    x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x
def stats(pos, cap=32):
    q = np.minimum((pos * 65536).astype(np.int64), 65535)
    key = (part1by2(q[:,0]) << np.uint64(2)) | (part1by2(q[:,1]) << np.uint64(1)) | part1by2(q[:,2])
    key.sort()
    alive = key
    sizes = []
    for L in range(1, 49):
        pre = alive >> np.uint64(48 - L)
        u, inv, cnt = np.unique(pre, return_inverse=True, return_counts=True)
        done = cnt <= cap
        sizes.append(cnt[done])
        alive = alive[~done[inv]]
        if alive.size == 0: break
    s = np.concatenate(sizes)
    return s.mean(), len(s), np.percentile(s,[5,50,95])
if __name__ == '__main__' and len(sys.argv) == 1:
  for nx in (128, 200, 256):
    g = (np.arange(nx) + 0.5) / nx
    pos = np.stack(np.meshgrid(g, g, g, indexing='ij'), -1).reshape(-1, 3)
    print('lattice', nx, stats(pos))
  rng = np.random.default_rng(1)
  print('random 2M', stats(rng.random((2_000_000, 3))))


def packed_stats(pos, cap=32, leaf=8, sup=256):
    """option group_pack: cells (maximal prefix nodes with <= leaf particles) packed greedily, in Morton order, into groups of <= cap
    inside every maximal prefix node with <= sup particles (tree.cu k_groups_packed)"""
    q = np.minimum((pos * 65536).astype(np.int64), 65535)
    key = (part1by2(q[:, 0]) << np.uint64(2)) | (part1by2(q[:, 1]) << np.uint64(1)) | part1by2(q[:, 2])
    key.sort()

    def node_id(limit):          # id of the maximal prefix node with <= limit particles that holds each particle
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
    cell, sup_id = node_id(leaf), node_id(sup)
    starts = np.flatnonzero(np.r_[True, cell[1:] != cell[:-1]])     # sorted order: runs of equal cell id are the cells
    counts = np.diff(np.r_[starts, key.size])
    sid = sup_id[starts]
    sizes, acc, prev = [], 0, None
    for c, s in zip(counts.tolist(), sid.tolist()):
        if acc and (s != prev or acc + c > cap):
            sizes.append(acc); acc = 0
        acc += c; prev = s
    sizes.append(acc)
    s = np.array(sizes)
    return s.mean(), len(s), np.percentile(s, [5, 50, 95])


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "packed":
    g = (np.arange(100) + 0.5) / 100
    pos = np.stack(np.meshgrid(g, g, g, indexing='ij'), -1).reshape(-1, 3)
    print('packed: lattice 100', packed_stats(pos), 'plain', stats(pos))
    rng = np.random.default_rng(1)
    r = rng.random((1_000_000, 3))
    print('packed: random 1M', packed_stats(r), 'plain', stats(r))

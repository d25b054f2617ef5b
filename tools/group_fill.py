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
for nx in (128, 200, 256):
    g = (np.arange(nx) + 0.5) / nx
    pos = np.stack(np.meshgrid(g, g, g, indexing='ij'), -1).reshape(-1, 3)
    print('lattice', nx, stats(pos))
rng = np.random.default_rng(1)
print('random 2M', stats(rng.random((2_000_000, 3))))

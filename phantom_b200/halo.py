"""Multi-GPU orchestration of the hot path: spatial domain decomposition + ghost-particle halo over NCCL.

Replaces, for one node with N GPUs (one process per GPU, torch.distributed/NCCL over NVLink), the reference's MPI stack
for this path:

  maketreeglobal / balancedomains   src/main/kdtree.F90:2044-2300, src/main/mpi_balance.F90:82
      -> `orb_boxes`: the same rule the reference applies globally -- split at the centre of mass along the longest
         axis, log2(N) levels -- giving one box per rank;
  send_cell / recv_cells / combine_cells  src/main/mpi_derivs.F90:197-522 (cell export, one round trip per pass)
      -> ghost particles: every rank receives the remote particles within radkern*hmax*margin of its box (stage 1,
         before the tree) and their post-density h, gradh, alpha (stage 2); ghosts enter as inactive particles, the
         owner alone accumulates its particles' sums, nothing is sent back;
  reduceall_mpi('min'|'max')  src/main/force.F90:848-852, src/main/dens.F90:546-549
      -> one packed all_reduce for dtcourant / dtforce / rhomax.

The product path is `DistSph` below: selection, packing, the grouped NCCL transfers, unpacking, the migration of particles that
leave their box and the reductions all run inside the library (phantom_b200/csrc/dist.cu, `sphgpu_dist_*` of include/sphgpu.h).
The numpy functions of this module restate the device kernels' rules (bisection, ownership, ghost selection) for the gloo tests,
which run the same protocol on CPU with the test suite's CPU restatement as the compute stand-in.
"""
import math
import numpy as np

RADKERN = {0: 2.0, 1: 3.0}


def orb_boxes(xyz, mass, nranks, box_lo, box_hi):
    """Recursive bisection of `box` at the centre of mass along the longest axis (kdtree.F90:2098-2160).
    xyz may be a sample of the particle set; returns (nranks, 6) boxes {lo, hi} tiling the input box."""
    boxes = [(np.array(box_lo, dtype=float), np.array(box_hi, dtype=float), np.arange(len(xyz)))]
    assert nranks & (nranks - 1) == 0, "number of ranks must be a power of two (log2 N bisection levels)"
    while len(boxes) < nranks:
        new = []
        for lo, hi, idx in boxes:
            axis = int(np.argmax(hi - lo))
            x = xyz[idx, axis]
            w = mass[idx] if np.ndim(mass) else np.full(len(idx), mass)
            pivot = float(np.sum(w * x) / np.sum(w)) if len(idx) else 0.5 * (lo[axis] + hi[axis])
            left = idx[x <= pivot]
            right = idx[x > pivot]
            hl, lr = hi.copy(), lo.copy()
            hl[axis] = pivot
            lr[axis] = pivot
            new.append((lo, hl, left))
            new.append((lr, hi, right))
        boxes = new
    return np.array([np.concatenate([lo, hi]) for lo, hi, _ in boxes])


def owner_of(xyz, boxes):
    """rank owning each position (half-open boxes: lo <= x < hi, the last box closed)"""
    owner = np.full(len(xyz), -1, dtype=np.int64)
    for r, b in enumerate(boxes):
        m = np.all((xyz >= b[:3]) & (xyz < b[3:]), axis=1) & (owner < 0)
        owner[m] = r
    # positions exactly on the global upper faces
    if np.any(owner < 0):
        for r, b in enumerate(boxes):
            m = np.all((xyz >= b[:3]) & (xyz <= b[3:]), axis=1) & (owner < 0)
            owner[m] = r
    return owner


def box_gap2(xyz, box, L, periodic):
    """squared minimum-image distance from points to a box (numpy restatement of k_halo_select)"""
    g2 = np.zeros(len(xyz))
    for k in range(3):
        x = xyz[:, k]
        g = np.maximum(0., np.maximum(box[k] - x, x - box[3 + k]))
        if periodic:
            for s in (-L[k], L[k]):
                g = np.minimum(g, np.maximum(0., np.maximum(box[k] - (x + s), (x + s) - box[3 + k])))
        g2 += g * g
    return g2


def select_ghosts_numpy(xyz, boxes, myrank, dhalo, L, periodic):
    """indices of owned particles each other rank needs"""
    return [np.nonzero(box_gap2(xyz, boxes[r], L, periodic) < dhalo * dhalo)[0] if r != myrank else np.zeros(0, dtype=np.int64)
            for r in range(len(boxes))]


def gather_padded(dist, torch, records):
    """all-gather of a different number of records per rank through padding to the largest count (numpy restatement of the
    device path k_gg_pack -> all_gather_into_tensor -> k_gg_unpack); returns (records of all ranks in rank order, counts, own offset)"""
    world, rank = dist.get_world_size(), dist.get_rank()
    n, w = records.shape
    cnt = torch.tensor([n], dtype=torch.int64)
    allc = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, cnt)
    counts = np.array([int(c[0]) for c in allc])
    stride = int(counts.max())
    padded = torch.zeros(stride * w, dtype=torch.float64)
    padded[: n * w] = torch.from_numpy(np.ascontiguousarray(records, dtype=np.float64).reshape(-1))
    recv = [torch.zeros(stride * w, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(recv, padded)
    glob = np.concatenate([recv[r].numpy().reshape(stride, w)[: counts[r]] for r in range(world)])
    return glob, counts, int(counts[:rank].sum())


class DistSph:
    """The decomposed particle set behind the C ABI (csrc/dist.cu): selection, packing, the NCCL exchanges, the migration of particles
    that leave their box and the reductions all run inside the library; this class only bootstraps the communicator (the 128-byte NCCL
    id travels over the existing torch.distributed group -- a Fortran host would MPI_Bcast it) and keeps the host-side slices."""

    def __init__(self, gpu, rank, world, boxes=None, domain=None, ids=None):
        import torch
        import torch.distributed as dist
        self.g, self.rank, self.world = gpu, rank, world
        uid = [gpu.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        gpu.dist_init(uid[0], world, rank)
        if ids is not None:
            gpu.dist_set_ids(ids)
        else:
            n = torch.tensor([gpu.npart_uploaded], dtype=torch.int64, device="cuda")
            alln = [torch.zeros_like(n) for _ in range(world)]
            dist.all_gather(alln, n)
            gpu.dist_set_ids(None, base=int(sum(int(a) for a in alln[:rank])))
        if boxes is not None:
            gpu.dist_set_boxes(boxes)
        else:
            gpu.dist_rebalance(domain)
        self.halo_bytes, self.nghost = 0, 0

    def derivs(self, icall=1, dt=0.0):
        sc = self.g.dist_derivs(icall, dt)
        st = self.g.dist_stats()
        self.halo_bytes, self.nghost, self.halo_rounds, self.ms = st["halo_bytes"], st["nghost"], st["rounds"], st["ms"]
        return sc

    def step(self, dt, tolv=1.e-2):
        return self.g.dist_step(dt, tolv)

    def energies(self):
        return self.g.dist_energies()

    def download_owned(self, params, fields=("xyzh", "vxyzu", "fxyzu")):
        """host copies of the owned particles' arrays and their global ids (the owned set changes as particles migrate)"""
        from .setups import Particles
        from . import api
        n = self.g.dist_nlocal()
        part = Particles(params, np.zeros((n, 4)))
        self.g.download(part, api.F_ALL)
        return part, self.g.dist_get_ids()

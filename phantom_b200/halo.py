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

The selection, packing and unpacking run on the device (phantom_b200/csrc/halo.cu); the exchange is an all-to-all-v
(`all_to_all_single`) issued directly on the library's device buffers.  With the gloo backend (CPU tests) the same
host logic runs against numpy restatements of the device kernels.
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


class _DevArray:
    """zero-copy view of a library-owned device buffer for torch (CUDA array interface)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class DistributedSph:
    """derivs on a domain-decomposed particle set: one instance per rank, SphGpu context underneath."""

    def __init__(self, gpu, boxes, rank, world, margin=1.15):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.g = gpu
        self.boxes = np.ascontiguousarray(boxes, dtype=np.float64)
        self.rank, self.world = rank, world
        self.margin = margin
        self.radkern = RADKERN[gpu.params.kernel]
        self.halo_bytes = 0
        self.nghost = 0
        self.nlocal = 0
        self._hu_prev = None
        # the context's compute stream becomes a blocking stream: implicitly ordered with the legacy default stream that torch and its
        # NCCL collectives work against, so pack -> all-to-all -> unpack needs no host synchronisation
        self.sync_free = False
        if torch.cuda.is_available() and torch.cuda.current_stream().cuda_stream == 0:
            gpu.set_option("legacy_stream", 1)
            self.sync_free = True

    def _alltoall(self, sendptr, rd, sendcounts, recvcounts, stage):
        torch, dist = self.torch, self.dist
        ntot_s, ntot_r = int(sendcounts.sum()), int(recvcounts.sum())
        recvptr = self.g.halo_recvbuf(ntot_r, rd)
        send = torch.as_tensor(_DevArray(sendptr, max(ntot_s, 1) * rd), device="cuda")[: ntot_s * rd]
        recv = torch.as_tensor(_DevArray(recvptr, max(ntot_r, 1) * rd), device="cuda")[: ntot_r * rd]
        dist.all_to_all_single(recv, send, [int(c) * rd for c in recvcounts], [int(c) * rd for c in sendcounts])
        if not self.sync_free:
            torch.cuda.synchronize()
        self.halo_bytes += 8 * rd * (ntot_s + ntot_r)
        return ntot_r

    def _gather_gravity_set(self):
        """all-gather of the owned particles' positions and h history: the input of the self-gravity pass (see include/sphgpu.h)"""
        torch, dist = self.torch, self.dist
        g = self.g
        nloc = torch.tensor([self.nlocal], dtype=torch.int64, device="cuda")
        allc = torch.empty(self.world, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(allc, nloc)
        counts = allc.cpu().numpy()
        stride = int(counts.max())
        sendptr, rd = g.gravity_gather_pack()
        recvptr = g.gravity_gather_recvbuf(self.world, stride)
        send = torch.as_tensor(_DevArray(sendptr, max(self.nlocal, 1) * rd), device="cuda")[: self.nlocal * rd]
        recv = torch.as_tensor(_DevArray(recvptr, self.world * stride * rd), device="cuda")
        padded = torch.zeros(stride * rd, dtype=torch.float64, device="cuda")
        padded[: self.nlocal * rd] = send
        dist.all_gather_into_tensor(recv, padded)
        torch.cuda.synchronize()
        self.halo_bytes += 8 * rd * stride * self.world
        g.gravity_gather_unpack(self.world, self.rank, stride, counts)

    def derivs(self, icall=1, dt=0.0):
        """tree + density + cons2prim + force with the two ghost exchanges; returns the reduced scalars"""
        torch, dist = self.torch, self.dist
        g = self.g
        self.halo_bytes = 0
        self.nlocal = int(g.npart_uploaded)
        import os, time
        timing = os.environ.get("SPHGPU_HALO_TIMING") and self.rank == 0
        marks = []

        def mark(name):
            if timing:
                torch.cuda.synchronize()
                marks.append((name, time.perf_counter()))
        mark("start")
        # halo width from the global hmax: one tiny all_reduce on the first call; afterwards the largest trial h of the previous
        # density pass (already reduced) -- the widening check below keeps this safe when h grows between steps
        if self._hu_prev is None:
            hm = torch.tensor([g.local_hmax()], dtype=torch.float64, device="cuda")
            dist.all_reduce(hm, op=dist.ReduceOp.MAX)
            self._hu_prev = float(hm[0])
        dhalo = self.radkern * self._hu_prev * self.margin
        self.halo_rounds = 0
        while True:
            self.halo_rounds += 1
            sendcounts = g.halo_select(self.world, self.rank, self.boxes, dhalo)
            sc = torch.from_numpy(sendcounts).cuda()
            rc = torch.empty_like(sc)
            dist.all_to_all_single(rc, sc)
            recvcounts = rc.cpu().numpy()
            ptr, rd = g.halo_pack(1)
            self.nghost = self._alltoall(ptr, rd, sendcounts, recvcounts, 1)
            g.halo_unpack(1, self.nghost)
            mark("halo1")
            g.build_tree_resident()
            mark("tree")
            sd = g.densityiterate_resident(1)
            mark("density")
            # widening round (the reference re-exports a cell whenever its h outgrows the search radius, dens.F90:343-365): if any
            # particle anywhere iterated with 2h beyond the halo the ghosts were selected with, restore h and repeat with a wider halo
            hu = torch.tensor([g.density_hmax_used(), g.density_hgrow()], dtype=torch.float64, device="cuda")
            dist.all_reduce(hu, op=dist.ReduceOp.MAX)
            hu = hu.cpu().numpy()
            self._hu_prev = float(hu[0])
            if self.radkern * float(hu[0]) <= dhalo or self.halo_rounds >= 8:
                break
            dhalo = self.radkern * float(hu[0]) * self.margin
            g.halo_restore_h()
        g.set_option("halo_hgrow", float(hu[1]))
        g.params.set_boundaries_to_active = 0
        ptr, rd = g.halo_pack(2)
        self._alltoall(ptr, rd, sendcounts, recvcounts, 2)
        g.halo_unpack(2, self.nghost)
        mark("halo2")
        g.cons2prim_resident()
        if g.params.gravity:
            self._gather_gravity_set()
        sf = g.force_resident(icall, dt)
        mark("c2p+force")
        red = torch.tensor([sf.dtcourant, sf.dtforce, -sd.rhomax], dtype=torch.float64, device="cuda")
        dist.all_reduce(red, op=dist.ReduceOp.MIN)
        red = red.cpu().numpy()                      # one device-to-host read for the three scalars
        sf.dtcourant, sf.dtforce, sf.rhomax = float(red[0]), float(red[1]), -float(red[2])
        sf.np, sf.nrhocalc, sf.npairs_density = sd.np, sd.nrhocalc, sd.npairs_density
        sf.actualmean, sf.maxactual, sf.trialmean, sf.nactualtot = sd.actualmean, sd.maxactual, sd.trialmean, sd.nactualtot
        mark("reduce")
        if timing:
            print("halo timing (ms): " + " ".join(f"{n}={1e3 * (t - marks[k][1]):.3f}" for k, (n, t) in enumerate(marks[1:])) +
                  f" total={1e3 * (marks[-1][1] - marks[0][1]):.3f} nghost={self.nghost}", flush=True)
        return sf


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

"""Worker for the world_size>1 tests (spawned by test_halo_gloo.py / test_gpu_multi.py).

backend=gloo : host logic of the multi-GPU path (ORB boxes, ownership, ghost selection radius, the two-stage ghost
               protocol) exercised on CPU with the oracle standing in for the device kernels, result compared with the
               oracle on the undivided particle set.
backend=nccl : the real thing -- SphGpu + DistSph (csrc/dist.cu behind the C ABI) on one GPU per rank, compared with the oracle on the whole set.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def global_problem(nx, mhd=False):
    from phantom_b200 import setups
    if nx < 0:              # turbulent box of (-nx)^3 particles (C2): the stepping test
        part = setups.setup_turb(nx=-nx)
        part.alphaind[:, 0] = 1.0
        return part
    if 500 <= nx < 1000:    # Sod shock tube (C1): boundary particles at the x ends (set_boundaries_to_active on the first pass), quintic kernel
        part = setups.setup_shock(nx=nx - 500)
        part.alphaind[:, 0] = 1.0
        return part
    if nx >= 1000:          # self-gravitating sphere of nx particles (C5): exercises the gathered gravity set
        part = setups.setup_random_sphere(n=nx)
        rng = setups.Ran2(-97531)
        part.vxyzu[:, :3] = 0.1 * (rng.draw(3 * part.npart).reshape(-1, 3) - 0.5)
        part.alphaind[:, 0] = 0.5
        return part
    part, _ = setups.setup_test_derivs(nx=nx, lattice="random", mhd=mhd)
    rng = setups.Ran2(-24358)
    part.xyzh[:, 3] *= (0.9 + 0.2 * rng.draw(part.npart))
    part.alphaind[:, 0] = 0.5
    return part


def take(part, idx):
    from phantom_b200.setups import Particles
    q = Particles(part.params.copy(), part.xyzh[idx])
    for k in ("vxyzu", "fxyzu", "fext", "Bevol", "iphase", "alphaind", "gradh", "divcurlv", "dvdx", "eos_vars"):
        getattr(q, k)[...] = getattr(part, k)[idx]
    return q


def main():
    backend, nx, outdir = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    from phantom_b200 import halo
    part = global_problem(nx)
    p = part.params
    L = np.array([p.xmax - p.xmin, p.ymax - p.ymin, p.zmax - p.zmin])
    boxes = halo.orb_boxes(part.xyzh[:, :3], p.massoftype[1], world, [p.xmin, p.ymin, p.zmin], [p.xmax, p.ymax, p.zmax])
    owner = halo.owner_of(part.xyzh[:, :3], boxes)
    assert np.all(owner >= 0)
    mine = np.nonzero(owner == rank)[0]
    loc = take(part, mine)

    if backend.startswith("nccl"):
        # the product path: everything behind the C ABI (csrc/dist.cu); only the NCCL id travels over torch.distributed
        torch.cuda.set_device(rank % torch.cuda.device_count())
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank % torch.cuda.device_count()))
        from phantom_b200.api import SphGpu
        g = SphGpu(loc.params.copy(), device=rank % torch.cuda.device_count())
        g.upload(loc)
        domain = [p.xmin, p.ymin, p.zmin, p.xmax, p.ymax, p.zmax]
        if backend == "nccl_rebalance":       # boxes from the library's own bisection; the particles start on the WRONG ranks (round robin)
            mine = np.arange(rank, part.npart, world)
            loc = take(part, mine)
            g.upload(loc)
            d = halo.DistSph(g, rank, world, domain=domain, ids=mine)
        else:
            d = halo.DistSph(g, rank, world, boxes=boxes, ids=mine)
        if backend == "nccl_step":
            nsteps = int(sys.argv[4])
            sc = d.derivs(1)
            dt = min(sc.dtcourant, sc.dtforce)
            hist, migrated = [], 0
            for it in range(nsteps):
                out = d.step(dt)
                migrated += g.dist_stats()["migrated"]
                e = d.energies()
                hist.append([dt, out.its, e.ekin, e.etherm, e.emag, e.epot, e.etot, e.xmom, e.ymom, e.zmom, e.mtot, float(e.np)])
                dt = min(out.dtcourant, out.dtforce, out.dterr)
            own, ids = d.download_owned(loc.params)
            np.savez(os.path.join(outdir, f"rank{rank}.npz"), idx=ids, xyzh=own.xyzh, vxyzu=own.vxyzu, hist=np.array(hist), migrated=migrated,
                     nghost=d.nghost)
        else:
            sc = d.derivs(1)
            own, ids = d.download_owned(loc.params)
            np.savez(os.path.join(outdir, f"rank{rank}.npz"), idx=ids, xyzh=own.xyzh, fxyzu=own.fxyzu, gradh=own.gradh, divcurlv=own.divcurlv,
                     dtcourant=sc.dtcourant, dtforce=sc.dtforce, nghost=d.nghost, poten=own.poten, nactualtot=sc.nactualtot, npairs_force=sc.npairs_force,
                     boxes=g.dist_get_boxes(world))
        g.dist_finalize()
        dist.destroy_process_group()
        return

    # ---------------- gloo: host logic with the oracle as the compute stand-in ----------------
    dist.init_process_group("gloo")
    if p.gravity:
        # bookkeeping of the gathered gravity set (halo.gather_padded = numpy restatement of k_gg_pack / k_gg_unpack + all-gather)
        rec = np.concatenate([loc.xyzh, mine[:, None].astype(np.float64)], axis=1)
        glob, counts, own_lo = halo.gather_padded(dist, torch, rec)
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), idx=mine, glob=glob, counts=counts, own_lo=own_lo)
        dist.destroy_process_group()
        return
    from oraclelib import Oracle
    radkern = halo.RADKERN[p.kernel]
    hm = torch.tensor([loc.xyzh[:, 3].max()], dtype=torch.float64)
    dist.all_reduce(hm, op=dist.ReduceOp.MAX)
    dhalo = radkern * float(hm[0]) * 1.15
    sel = halo.select_ghosts_numpy(loc.xyzh[:, :3], boxes, rank, dhalo, L, bool(p.periodic))

    def exchange(fields):
        send = [torch.from_numpy(np.ascontiguousarray(np.concatenate([f[s].reshape(len(s), int(np.prod(f.shape[1:])) if f.ndim > 1 else 1) for f in fields], axis=1)))
                for s in sel]
        cnt = torch.tensor([len(s) for s in sel], dtype=torch.int64)
        rcnt = torch.empty_like(cnt)
        dist.all_to_all_single(rcnt, cnt)
        width = send[0].shape[1]
        sflat = torch.cat([x.reshape(-1) for x in send])
        rflat = torch.empty(int(rcnt.sum()) * width, dtype=torch.float64)
        dist.all_to_all_single(rflat, sflat, [int(c) * width for c in rcnt], [int(c) * width for c in cnt])
        return rflat.numpy().reshape(-1, width)

    nvu = p.maxvxyzu
    g1 = exchange([loc.xyzh, loc.vxyzu, loc.fxyzu[:, :3] + loc.fext, loc.Bevol, loc.iphase.astype(np.float64)])
    ng = len(g1)
    from phantom_b200.setups import Particles
    both = Particles(p.copy(), np.concatenate([loc.xyzh, g1[:, :4]]))
    nl = loc.npart
    for k in ("vxyzu", "fxyzu", "fext", "Bevol", "alphaind", "gradh"):
        getattr(both, k)[:nl] = getattr(loc, k)
    both.vxyzu[nl:] = g1[:, 4:4 + nvu]
    both.fxyzu[nl:, :3] = g1[:, 4 + nvu:7 + nvu]
    both.Bevol[nl:] = g1[:, 7 + nvu:11 + nvu]
    both.iphase[:nl] = loc.iphase
    both.iphase[nl:] = -np.abs(g1[:, 11 + nvu]).astype(np.int8)      # ghosts: inactive, neighbour only
    o = Oracle(both.params)
    o.build_tree(both)
    o.densityiterate(both, 1)
    both.params.set_boundaries_to_active = 0
    o.set_params(both.params)
    # stage 2: owners send the post-density h, gradh, alpha of the same ghost sets
    loc_h = both.xyzh[:nl, 3:4]
    g2 = exchange([loc_h, both.gradh[:nl].astype(np.float64), both.alphaind[:nl, 0:1].astype(np.float64)])
    both.xyzh[nl:, 3] = g2[:, 0]
    both.gradh[nl:, 0] = g2[:, 1].astype(np.float32)
    both.alphaind[nl:, 0] = g2[:, 2].astype(np.float32)
    o.build_tree(both)            # refit of hmax with the ghosts' new h (the oracle has no refit: rebuild)
    o.cons2prim(both)
    sf = o.force(both, 1)
    red = torch.tensor([sf.dtcourant, sf.dtforce], dtype=torch.float64)
    dist.all_reduce(red, op=dist.ReduceOp.MIN)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), idx=mine, xyzh=both.xyzh[:nl], fxyzu=both.fxyzu[:nl], gradh=both.gradh[:nl],
             divcurlv=both.divcurlv[:nl], dtcourant=float(red[0]), dtforce=float(red[1]), nghost=ng)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

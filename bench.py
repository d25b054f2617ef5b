#!/usr/bin/env python
"""bench.py -- particle-updates/s of the SPH hot path (tree + density + cons2prim + force) on B200.

One "step" = one derivs(icall=1) over the whole synthetic particle set (BASELINE.json metric).
  value : device-resident throughput, timed on the device with CUDA events, max over ranks
  e2e   : the same step through the reference-facing C-ABI call sphgpu_derivs with HOST (pinned) buffers,
          host->device and device->host copies inside the timed region
  roofline     : the dominant pair kernel against the FP64 pipe peak measured on this device (DFMA microbenchmark)
  cpu_baseline : the CPU oracle ("port" of the reference algorithm; the reference is Fortran and cannot be built)
`--impl reference` times that CPU implementation alone on the host cores with the same config/metric.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def flop_constants():
    txt = open(os.path.join(ROOT, "phantom_b200", "csrc", "roofline_constants.h")).read()
    return {k: int(v) for k, v in re.findall(r"#define\s+(\w+)\s+(\d+)", txt)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons streamed DURING the timed region (B200_PROFILING.md recipe, -lms 100)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def start(self):
        time.sleep(0.35)      # let the first samples land before the timed region starts

    def stop(self):
        self.samples = []
        if self.proc is None:
            return
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) >= 7:
                self.samples.append(f)

    def summary(self):
        if not getattr(self, "samples", None):
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": sorted(reasons),
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(self.samples)}


def bind_to_gpu_numa_node(local):
    """pin this rank's host threads to the CPUs NVML names as local to its GPU, so pinned buffers and the DMA stay on that socket"""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else local
        hdl = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(hdl, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as e:   # not fatal: the bench still runs, just without the binding
        print(f"bench.py: no NUMA binding ({type(e).__name__}: {e})", file=sys.stderr)


def pinned_like(part):
    """re-home the particle arrays in pinned host memory so the e2e copies are the DMA path a resident host code would use"""
    import torch
    for k, v in list(part.__dict__.items()):
        if isinstance(v, np.ndarray) and v.size > 0:
            t = torch.empty(v.shape, dtype=torch.from_numpy(v).dtype, pin_memory=True)
            a = t.numpy()
            a[...] = v
            setattr(part, k, a)
            part.__dict__.setdefault("_pins", []).append(t)
    return part


def make_workload(args, rank=0, world=1):
    from phantom_b200 import setups
    if world == 1:
        part = setups.setup_turb(nx=args.nx, positions=getattr(args, "positions", "lattice"))
        pos = "" if getattr(args, "positions", "lattice") == "lattice" else ", uniformly random positions"
        return part, None, f"turb: isothermal periodic box, {args.nx}^3 = {part.npart} particles{pos}, cubic kernel, hydro+AV (Cullen-Dehnen), all active"
    part, boxes = setups.setup_turb_block(args.nx, world, rank)
    bd = setups.block_dims(world)
    return part, boxes, (f"turb (weak scaling): periodic box of {bd[0]}x{bd[1]}x{bd[2]} blocks, {args.nx}^3 = {part.npart} particles per GPU, "
                         f"{part.npart * world} in total, cubic kernel, hydro+AV (Cullen-Dehnen), all active")


def parity_block(SphGpu, pin, pcpu, sdo, sfo, device):
    """GPU (C-ABI sphgpu_derivs on a fresh copy of the input) against the CPU arm's result on the same input, every particle."""
    pg = pin.copy()
    g = SphGpu(pg.params.copy(), device=device)
    sg = g.derivs(pg, 1)

    def relmax(a, b):
        sc = float(np.sqrt(np.mean(b.astype(np.float64) ** 2))) + 1e-300
        return float(np.max(np.abs(a - b) / (np.abs(b) + sc)))
    out = {"particles": int(pin.npart),
           "max_rel_h": float(np.max(np.abs(pg.xyzh[:, 3] - pcpu.xyzh[:, 3]) / pcpu.xyzh[:, 3])),
           "max_rel_f": relmax(pg.fxyzu[:, :3], pcpu.fxyzu[:, :3]),
           "nactualtot_equal": bool(sg.nactualtot == sdo.nactualtot and sg.maxactual == sdo.maxactual),
           "npairs_force_equal": bool(sg.npairs_force == sfo.npairs_force),
           "nactualtot": int(sg.nactualtot), "npairs_force": int(sg.npairs_force),
           "dtcourant_rel": float(abs(sg.dtcourant - sfo.dtcourant) / sfo.dtcourant),
           "tolerances": {"h": 1e-10, "f": 1e-8}}
    if pin.params.maxvxyzu == 4:
        out["max_rel_dudt"] = relmax(pg.fxyzu[:, 3], pcpu.fxyzu[:, 3])
    if pin.params.mhd:
        out["max_rel_dBdt"] = relmax(pg.dBevol, pcpu.dBevol)
    out["ok"] = bool(out["max_rel_h"] < 1e-10 and out["max_rel_f"] < 1e-8 and out.get("max_rel_dudt", 0.) < 1e-8 and out.get("max_rel_dBdt", 0.) < 1e-8
                     and out["nactualtot_equal"] and out["npairs_force_equal"])
    return out


def run_reference(args, rank, world):
    """CPU arm: the oracle (port of the reference algorithm) on all host threads; rank 0 only."""
    if rank != 0:
        return
    from oraclelib import Oracle
    part, _, wl = make_workload(args)
    o = Oracle(part.params)
    # all the host threads the box offers: torchrun exports OMP_NUM_THREADS=1 to its workers, which would otherwise leave the CPU arm
    # on one core whenever N > 1
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if o.max_threads() < ncpu and (world > 1 or "TORCHELASTIC_RUN_ID" in os.environ):
        o.set_threads(ncpu)
    threads = o.max_threads()
    # bounded sample: one derivs at full size is timed first (it doubles as a warm-up step); the box is halved only if
    # (steps + warmup) such steps would not fit in ~150 s (cost is linear in N)
    nx = nx_asked = args.nx
    t1 = time.perf_counter()
    o.derivs(part)
    t_step = time.perf_counter() - t1
    done_warm = 1
    while nx > 32 and t_step * (args.steps + max(args.warmup - 1, 0)) > 150.0:
        nx //= 2
        t_step /= 8.0
    if nx != args.nx:
        args.nx = nx
        part, _, _ = make_workload(args)
        o = Oracle(part.params)
        done_warm = 0
    for _ in range(max(args.warmup - done_warm, 0)):
        o.derivs(part)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.derivs(part)
    dt = time.perf_counter() - t0
    val = part.npart * args.steps / dt
    sample = f"{args.steps} full derivs on {nx}^3 = {part.npart} particles"
    line = {
        "impl": "reference", "metric": "particle-updates/s (tree+density+cons2prim+force)", "value": val, "unit": "particle-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict({"workload": wl}, **({"cpu_sample_nx": nx} if nx != nx_asked else {})),
        "cpu_baseline": {"value": val, "unit": "particle-updates/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: everything else written to fd 1 by libraries (NCCL prints its version there) goes to stderr"""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--nx", type=int, default=128, help="lattice points per axis of the turb box (BASELINE configs[1]: 128)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--max-cell", type=int, default=0)
    ap.add_argument("--max-leaf", type=int, default=0)
    ap.add_argument("--hilbert", action="store_true", help="A/B: Hilbert instead of Morton particle order")
    ap.add_argument("--positions", default="lattice", choices=["lattice", "random"], help="N=1: lattice (BASELINE config) or uniformly random positions")
    ap.add_argument("--evolve", type=int, default=0, help="N=1: run this many leapfrog steps first, then time derivs from the PREDICTED h of the next step "
                                                          "(so the h-rho iteration does real work)")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip the 8 M particles per GPU weak-scaling block")
    ap.add_argument("--extra-nx", type=int, default=200)
    ap.add_argument("--extra-steps", type=int, default=5)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        bind_to_gpu_numa_node(local)     # (N = 1 keeps every host core: the CPU baseline of the same run uses them all)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from phantom_b200.api import SphGpu, F_ALL, F_XYZH
    from phantom_b200.halo import DistSph
    part, boxes, wl = make_workload(args, rank, world)
    n = part.npart
    g = SphGpu(part.params.copy(), device=local)
    if args.max_cell:
        g.set_option("max_cell", args.max_cell)
    if args.max_leaf:
        g.set_option("max_leaf", args.max_leaf)
    if args.hilbert:
        g.set_option("hilbert", 1)
    fp64_peak = g.measure_fp64_peak()
    copy_bw = g.measure_copy_bw()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm ----------------
    g.upload(part)
    # N > 1: the whole exchange runs behind the C ABI (csrc/dist.cu: selection, packing, grouped ncclSend/ncclRecv, unpacking and the
    # reductions are queued on the library's stream); torch.distributed only carries the 128-byte NCCL id
    dsph = DistSph(g, rank, world, boxes=boxes) if world > 1 else None

    state = "device-resident, derivs(icall=1) repeated on the same state"
    pred = None
    if args.evolve > 0 and world == 1:
        # an evolved, disordered state: K leapfrog steps on the device, then every timed derivs starts from the h that predict_sph would
        # hand it for the next step (step_leapfrog.f90:332: h <- h - dt dh/drho rho div v = h (1 + dt div v / 3)), re-uploaded outside
        # the timed regions (the timings are CUDA events inside the library)
        sc0 = g.derivs_resident(1)
        dt = min(sc0.dtcourant, sc0.dtforce)
        for _ in range(args.evolve):
            out = g.step_resident(dt)
            dt = min(out.dtcourant, out.dtforce, out.dterr)
        g.download(part)
        pred = part.copy()
        pred.xyzh[:, 3] = part.xyzh[:, 3] * (1. + dt * part.divcurlv[:, 0].astype(np.float64) / 3.)
        pred = pinned_like(pred)
        state = f"evolved for {args.evolve} leapfrog steps on the device; every timed derivs(1) starts from the h predicted for the next step (dt = {dt:.3e})"

    def step():
        if pred is not None:
            g.upload(pred, F_XYZH)
        return dsph.derivs(1) if dsph else g.derivs_resident(1)

    for _ in range(args.warmup):
        sc = step()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = g.launch_count()
    barrier()
    t_dev = 0.0
    phases = dict(tree=0.0, dens=0.0, cons2prim=0.0, force=0.0)
    kern = dict(density=0.0, force=0.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sc = step()
        if world == 1:
            tm = g.timings_ms()
            t_dev += sum(tm.values())
            for k in phases:
                phases[k] += tm[k]
        else:
            t_dev += dsph.ms                  # CUDA events on the library's stream around the whole call (kernels + NCCL transfers)
        kt = g.kernel_timings_ms()
        for k in kern:
            kern[k] += kt[k]
    barrier()
    t_wall = time.perf_counter() - t0
    launches = g.launch_count() - l0
    sampler.stop()
    # max over ranks of the device time
    tt = torch.tensor([t_dev, t_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev_max, t_wall_max = float(tt[0]) * 1e-3, float(tt[1])
    value = n * world * args.steps / t_dev_max
    halo_info = None
    if dsph:
        hb = torch.tensor([float(dsph.nghost), float(dsph.halo_bytes)], dtype=torch.float64, device="cuda")
        dist.all_reduce(hb, op=dist.ReduceOp.MAX)
        halo_info = {"ghosts_per_gpu_max": int(hb[0]), "halo_bytes_per_step_per_gpu_max": int(hb[1]), "exchanges_per_step": 2,
                     "collective": "grouped ncclSend/ncclRecv of fixed-capacity ghost blocks on the library's stream (no count exchange, no host "
                                   "round trip per exchange) + 2 all-reduces per step, all behind the C ABI (sphgpu_dist_derivs)"}

    # ---------------- end-to-end arm: the literal C-ABI call with host (pinned) buffers ----------------
    part_e2e = pinned_like(part.copy())
    g2 = SphGpu(part.params.copy(), device=local)
    if args.max_cell:
        g2.set_option("max_cell", args.max_cell)
    if args.max_leaf:
        g2.set_option("max_leaf", args.max_leaf)
    nvu, ng = part.params.maxvxyzu, part.params.ngradh
    # N > 1: the same arrays the literal sphgpu_derivs call moves for an all-active hydro set (its inputs in, every array the passes
    # write out): xyzh iphase vxyzu fxyzu fext alphaind -> ; -> xyzh gradh dvdx eos_vars alphaind fxyzu divcurlv
    from phantom_b200 import api as A
    in_mask = A.F_XYZH | A.F_IPHASE | A.F_VXYZU | A.F_FXYZU | A.F_FEXT | A.F_ALPHAIND
    out_mask = A.F_XYZH | A.F_GRADH | A.F_DVDX | A.F_EOSVARS | A.F_ALPHAIND | A.F_FXYZU | A.F_DIVCURLV
    h2d = n * (4 * 8 + 1 + nvu * 8 * 2 + 3 * 8 + 3 * 4)
    d2h = n * (4 * 8 + ng * 4 + 9 * 4 + 7 * 8 + 3 * 4 + nvu * 8 + 4)
    if world > 1:
        g2.upload(part_e2e)          # every array once, outside the timed region (sizes the device buffers)
    dsph2 = DistSph(g2, rank, world, boxes=boxes) if world > 1 else None

    def step_e2e():
        if dsph2:
            g2.upload(part_e2e, in_mask)
            dsph2.derivs(1)
            g2.download(part_e2e, out_mask)
        else:
            g2.derivs(part_e2e, 1)

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = n * world * args.steps / float(te[0])
    if not dsph2:
        h2d, d2h = g2.copy_bytes()       # what the last sphgpu_derivs call moved over PCIe, counted by the library from the arrays it copies

    # ---------------- roofline of the dominant pair kernels ----------------
    fc = flop_constants()
    iso = bool(part.params.isothermal)
    # N > 1: the scalars are sums over the ranks, the kernel times are per GPU: the roofline is that of ONE GPU's launch
    npd, npf, nrc = sc.npairs_density / world, sc.npairs_force / world, sc.nrhocalc / world
    f_dens = fc["FLOP_DENS_PAIR_HYDRO"] * npd + fc["FLOP_DENS_EPILOGUE"] * nrc
    f_force = (fc["FLOP_FORCE_PAIR_ISOTHERMAL"] if iso else fc["FLOP_FORCE_PAIR_ADIABATIC"]) * npf + fc["FLOP_FORCE_EPILOGUE"] * n
    ms_d, ms_f = kern["density"] / args.steps, kern["force"] / args.steps
    passes = {
        "density": {"flops_per_launch": f_dens, "ms": ms_d, "achieved_tflops": f_dens / (ms_d * 1e-3) / 1e12, "pairs": int(npd),
                    "its_mean": sc.nrhocalc / max(sc.np, 1)},
        "force": {"flops_per_launch": f_force, "ms": ms_f, "achieved_tflops": f_force / (ms_f * 1e-3) / 1e12, "pairs": int(npf)},
    }
    for v in passes.values():
        v["frac_fp64"] = v["achieved_tflops"] / fp64_peak
    dom = "density" if ms_d >= ms_f else "force"
    bytes_step = n * (fc["BYTES_TREE_PER_PARTICLE"] + fc["BYTES_DENS_PER_PARTICLE"] + fc["BYTES_C2P_PER_PARTICLE"] + fc["BYTES_FORCE_PER_PARTICLE"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this same command
    # (profiles/ncu_traffic.json, written by tools/make_profile_summaries.py); null when no capture covers this workload
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if tr.get("workload_particles") == n:
            traffic = tr["kernels"].get("k_density" if dom == "density" else "k_force")
    except Exception:
        pass
    roofline = {
        "bound": "fp64", "kernel": "k_density" if dom == "density" else "k_force",
        "achieved": passes[dom]["achieved_tflops"], "peak": fp64_peak, "unit": "TFLOP/s", "frac": passes[dom]["frac_fp64"],
        "peak_source": "DFMA microbenchmark measured live on this device (MEASURED_PEAKS.json has no FP64 figure)",
        "traffic": traffic,
        "hbm_view": {"algorithmic_bytes_per_step": bytes_step, "achieved_gbs": bytes_step / (t_dev_max / args.steps) / 1e9, "peak_gbs": hbm_peak,
                     "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback", "frac": bytes_step / (t_dev_max / args.steps) / 1e9 / hbm_peak,
                     "copy_bw_live_gbs": copy_bw},
        "passes": passes,
    }

    line = {
        "metric": "particle-updates/s (tree+density+cons2prim+force)", "value": value, "unit": "particle-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl, "per_gpu_particles": n, "parallelism": "single" if world == 1 else f"spatial domain decomposition over {world} GPUs, ghost-particle halo (NCCL all-to-all-v), 1 process per GPU",
                   "timing": "CUDA events on the library stream" if world == 1 else
                   "CUDA events on the library stream around each sphgpu_dist_derivs (kernels and NCCL transfers share that stream), max over ranks",
                   "l2": "inputs_exceed_l2 (working set ~%.1f GB per step)" % (n * 600 / 1e9), "state": state},
        "phases_ms": {k: v / args.steps for k, v in phases.items()},
        "wall_ms_per_step": 1e3 * t_wall_max / args.steps,
        "e2e": {"value": e2e_val, "unit": "particle-updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "call": "sphgpu_derivs (literal C-ABI, pinned host buffers)" if world == 1 else "sphgpu_upload(inputs) + sphgpu_dist_derivs (halo exchanges) + sphgpu_download(outputs), pinned host buffers"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": roofline,
        "neighbours": {"mean": sc.actualmean, "max": sc.maxactual, "trial_mean": sc.trialmean},
        "halo": halo_info,
    }

    # ---------------- the size BASELINE's 80 % target is stated for: 200^3 = 8 M particles per GPU (64 M on 8 GPUs) ----------------
    if args.extras:
        del g2, dsph2, part_e2e
        import copy
        a2 = copy.copy(args)
        a2.nx = args.extra_nx
        pb, bb, wlb = make_workload(a2, rank, world)
        gb = SphGpu(pb.params.copy(), device=local)
        gb.upload(pb)
        db = DistSph(gb, rank, world, boxes=bb) if world > 1 else None
        tb = 0.0
        for it in range(2 + args.extra_steps):
            if db:
                db.derivs(1)
                ms = db.ms
            else:
                gb.derivs_resident(1)
                ms = sum(gb.timings_ms().values())
            if it >= 2:
                tb += ms
        tbt = torch.tensor([tb], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tbt, op=dist.ReduceOp.MAX)
        line["weak_scaling_8M_per_gpu"] = {"workload": wlb, "particles_total": int(pb.npart * world), "steps": args.extra_steps, "warmup": 2,
                                           "ms_per_step": float(tbt[0]) / args.extra_steps,
                                           "value": pb.npart * world * args.extra_steps / (float(tbt[0]) * 1e-3), "unit": "particle-updates/s",
                                           "note": "efficiency at N GPUs = value(N) / (N x value(1)) of this block across the per-N runs"}
        if db:
            gb.dist_finalize()
        del gb, db, pb

    # ---------------- N = 1 extras: an evolved (disordered) state, and parity on an MHD configuration ----------------
    if args.extras and world == 1 and args.evolve == 0:
        ge = SphGpu(part.params.copy(), device=local)
        pe = part.copy()
        ge.upload(pe)
        sc0 = ge.derivs_resident(1)
        dte = min(sc0.dtcourant, sc0.dtforce)
        for _ in range(20):
            oe = ge.step_resident(dte)
            dte = min(oe.dtcourant, oe.dtforce, oe.dterr)
        ge.download(pe)
        pred = pe.copy()
        pred.xyzh[:, 3] = pe.xyzh[:, 3] * (1. + dte * pe.divcurlv[:, 0].astype(np.float64) / 3.)
        te, its = 0.0, 0.0
        for it in range(2 + 5):
            ge.upload(pred, F_XYZH)
            sce = ge.derivs_resident(1)
            if it >= 2:
                te += sum(ge.timings_ms().values())
                its = sce.nrhocalc / max(sce.np, 1)
        line["evolved_state"] = {"what": "the same box after 20 leapfrog steps on the device (Mach 5), every timed derivs(1) starting from the h "
                                         "predicted for the next step, so that the h-rho iteration does real work",
                                 "ms_per_step": te / 5, "value": n * 5 / (te * 1e-3), "unit": "particle-updates/s", "its_mean": its,
                                 "candidates_per_group": sce.trialmean, "neighbours_mean": sce.actualmean}
        del ge
        # ---- the same evolved box with individual timesteps (BASELINE configs[1]: "IND_TIMESTEPS on/off both reported"): one dtmax of
        # substeps on the device -- set_active_particles + step() per smallest timestep, the bin bookkeeping of evolve.F90 on the host
        try:
            import time as _time

            def _ind_cycle(refcompat):
                pi_ = pe.copy()
                pi_.params.ind_timesteps = 1
                dtmax_i = 8. * dte
                pi_.params.dtmax = dtmax_i
                gi = SphGpu(pi_.params.copy(), device=local)
                if refcompat is not None:
                    gi.set_option("refcompat_hmax", float(refcompat))
                gi.upload(pi_)
                gi.set_timestep_bins(0, 0, 0)
                sci = gi.derivs_resident(1)
                nbinmax = int(sci.nbinmaxnew)
                gi.init_step_resident(0., dtmax_i, nbinmax)
                istepfrac, ti, nsub, nact_tot, nb0 = 0, 0., 0, 0, nbinmax
                torch.cuda.synchronize()
                t0 = _time.perf_counter()
                tsub = []
                while nsub < 256:
                    dti = dtmax_i / 2 ** nbinmax
                    istepfrac += 1
                    ts0 = _time.perf_counter()
                    nactive, _nalive = gi.set_active_particles_resident(nbinmax, istepfrac)
                    outi = gi.step_ind_resident(ti, dti, dtmax_i)
                    tsub.append((_time.perf_counter() - ts0) * 1e3)
                    nsub += 1; nact_tot += int(nactive)
                    ti = istepfrac / 2. ** nbinmax * dtmax_i
                    nbnew = int(outi.scalars.nbinmaxnew)
                    if nbnew != nbinmax:                               # change_nbinmax (utils_indtimesteps.f90:186-222)
                        istepfrac = istepfrac // 2 ** (nbinmax - nbnew) if nbnew < nbinmax else istepfrac * 2 ** (nbnew - nbinmax)
                        nbinmax = nbnew
                    if istepfrac == 2 ** nbinmax:
                        break
                torch.cuda.synchronize()
                wall_i = _time.perf_counter() - t0
                del gi
                return dict(nbinmax_start=nb0, nbinmax_end=nbinmax, substeps=nsub, active_updates=nact_tot, mean_active_fraction=nact_tot / max(nsub * n, 1),
                            ms_total=wall_i * 1e3, ms_per_substep=wall_i * 1e3 / max(nsub, 1), ms_per_substep_median=float(np.median(tsub)),
                            ms_first_substeps=[round(t, 2) for t in tsub[:3]], value=nact_tot / wall_i, unit="active-particle-updates/s",
                            synchronised=bool(istepfrac == 2 ** nbinmax))
            blk = _ind_cycle(None)
            exact = _ind_cycle(0)
            blk["what"] = ("the evolved box stepped through one dtmax = 8 x its global timestep with individual timestep bins "
                           "(sphgpu_set_active_particles_resident + sphgpu_step_ind_resident per smallest timestep, host wall clock around the loop); "
                           "default = neighbour sets pruned as the reference's tree walk prunes them (option refcompat_hmax, needs the reference-"
                           "topology tree every substep)")
            blk["global_equivalent_ms"] = blk["substeps"] * (te / 5)
            blk["exact_neighbour_sets"] = {"ms_per_substep": exact["ms_per_substep"], "ms_per_substep_median": exact["ms_per_substep_median"],
                                           "value": exact["value"], "substeps": exact["substeps"]}
            blk["note"] = "one cycle on a fresh context: the first substeps carry the one-time allocations of the reference tree (ms_first_substeps)"
            line["ind_timesteps"] = blk
        except Exception as e:      # an extra, never the reason for a missing bench line
            line["ind_timesteps"] = {"error": str(e)[:200]}
        if not args.no_cpu_baseline:
            from phantom_b200 import setups
            from oraclelib import Oracle
            pm = setups.setup_orstang(nx=48)
            pm.alphaind[:, 0] = 1.0
            pmc = pm.copy()
            om = Oracle(pmc.params)
            sdm, sfm = om.derivs(pmc)
            line["parity_mhd"] = dict(parity_block(SphGpu, pm, pmc, sdm, sfm, local), workload="Orszag-Tang vortex, MHD + div-B cleaning (BASELINE configs[2] at reduced size)")

    # ---------------- CPU baseline (oracle port) beside it: rank 0, N=1 ----------------
    if world == 1 and not args.no_cpu_baseline and rank == 0:
        from oraclelib import Oracle
        pc = part.copy()
        o = Oracle(pc.params)
        threads = o.max_threads()
        nxc = args.nx
        if 9.0e-6 * pc.npart * 8.0 / max(threads, 1) > 40.0:     # keep to ~10-30 s of CPU work
            nxc = args.nx // 2
            import copy
            a2 = copy.copy(args)
            a2.nx = nxc
            pc, _, _ = make_workload(a2)
            o = Oracle(pc.params)
        pin = pc.copy()
        t0 = time.perf_counter()
        sdo, sfo = o.derivs(pc)
        tc = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": pc.npart / tc, "unit": "particle-updates/s", "cores": threads, "kind": "port",
                                "sample": f"1 full derivs on {nxc}^3 = {pc.npart} particles ({tc:.1f} s)"}
        # parity of the two arms on the SAME input (outside every timed region): the C-ABI derivs on a fresh copy of the CPU arm's
        # initial state against the CPU arm's result, every particle, north_star tolerances (h 1e-10, a and du/dt 1e-8 relative)
        line["parity"] = parity_block(SphGpu, pin, pc, sdo, sfo, local)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and line.get("parity_mhd") and not line["parity_mhd"]["ok"]:
        raise SystemExit("bench.py: GPU and CPU arms disagree on the MHD configuration: %s" % json.dumps(line["parity_mhd"]))
    if rank == 0 and line.get("parity") and not line["parity"]["ok"]:
        raise SystemExit("bench.py: GPU and CPU arms disagree beyond the north_star tolerances: %s" % json.dumps(line["parity"]))


if __name__ == "__main__":
    main()

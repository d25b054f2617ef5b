"""2-GPU run of the real multi-GPU path (NCCL halo exchange) against the undivided oracle result; skipped with < 2 GPUs."""
import tempfile

import pytest

pytestmark = pytest.mark.gpu


def test_two_gpu_halo_exchange_matches_undivided():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from test_halo_gloo import launch, check_against_undivided
    with tempfile.TemporaryDirectory() as d:
        launch("nccl", 2, 20, d, 29611)
        check_against_undivided(d, 2, 20)


def test_two_gpu_selfgravity_matches_undivided():
    """self-gravity over 2 GPUs (gathered particle set, same tree on every rank) against the oracle on the undivided sphere"""
    import os
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from test_halo_gloo import launch
    from halo_worker import global_problem
    from oraclelib import Oracle
    n = 4000
    with tempfile.TemporaryDirectory() as d:
        launch("nccl", 2, n, d, 29613)
        ref = global_problem(n)
        sd, sf = Oracle(ref.params).derivs(ref)
        fs = np.sqrt(np.mean(ref.fxyzu[:, :3] ** 2))
        seen = np.zeros(n, dtype=int)
        for r in range(2):
            o = np.load(os.path.join(d, f"rank{r}.npz"))
            idx = o["idx"]
            seen[idx] += 1
            assert np.max(np.abs(o["xyzh"][:, 3] - ref.xyzh[idx, 3]) / ref.xyzh[idx, 3]) < 1e-10
            assert np.max(np.abs(o["fxyzu"][:, :3] - ref.fxyzu[idx, :3])) < 1e-8 * fs
            assert np.max(np.abs(o["poten"] - ref.poten[idx])) <= 3e-7 * np.max(np.abs(ref.poten))
            assert abs(o["dtforce"] - sf.dtforce) < 1e-8 * sf.dtforce
        assert np.all(seen == 1)


def test_two_gpu_rebalance_from_wrong_ranks_matches_undivided():
    """particles dealt round-robin to the ranks: the library's own bisection + migration must sort them out, then derivs == oracle"""
    import os
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from test_halo_gloo import launch, check_against_undivided
    with tempfile.TemporaryDirectory() as d:
        launch("nccl_rebalance", 2, 20, d, 29615)
        check_against_undivided(d, 2, 20)
        b = np.load(os.path.join(d, "rank0.npz"))["boxes"]
        assert abs(np.sum(np.prod(b[:, 3:] - b[:, :3], axis=1)) - 1.0) < 1e-12          # the boxes tile the unit box
        n0, n1 = [len(np.load(os.path.join(d, f"rank{r}.npz"))["idx"]) for r in range(2)]
        assert abs(n0 - n1) < 0.2 * (n0 + n1)                                              # centre-of-mass bisection balances the load


def test_two_gpu_step_with_migration_tracks_undivided_oracle():
    """30 leapfrog steps of the turbulent box on 2 GPUs (sphgpu_dist_step: migration + ghost exchanges inside every derivs) against the
    numpy restatement of the step driving the oracle on the UNDIVIDED set: energies and momentum to 1e-10 (north_star), same dt sequence"""
    import os
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from test_halo_gloo import launch
    from halo_worker import global_problem
    from oraclelib import Oracle
    import steplib
    nsteps = 30
    with tempfile.TemporaryDirectory() as d:
        launch("nccl_step", 2, -16, d, 29617, nsteps)
        outs = [np.load(os.path.join(d, f"rank{r}.npz")) for r in range(2)]
    ref = global_problem(-16)
    o = Oracle(ref.params)
    sc = steplib.oracle_derivs(o, ref, 1)
    dt = min(sc.dtcourant, sc.dtforce)
    hist = outs[0]["hist"]
    assert np.array_equal(hist, outs[1]["hist"])                       # every rank sees the same reduced numbers
    e0 = None
    for it in range(nsteps):
        assert abs(hist[it, 0] - dt) <= 1e-9 * dt, (it, hist[it, 0], dt)
        sc, dterr, errmax, its = steplib.step_leapfrog(o, ref, dt)
        assert int(hist[it, 1]) == its
        e = steplib.energies(ref)
        if e0 is None:
            e0 = e
            escale = abs(e["ekin"]) + abs(e["etherm"]) + abs(e["emag"]) + abs(e["epot"])
            pscale = np.sqrt(2. * max(e["ekin"], 1e-300) * e["mtot"])
        assert abs(hist[it, 6] - e["etot"]) <= 1e-10 * escale, (it, hist[it, 6], e["etot"])
        assert abs(hist[it, 2] - e["ekin"]) <= 1e-10 * escale
        assert np.max(np.abs(hist[it, 7:10] - e["mom"])) <= 1e-10 * pscale
        assert abs(hist[it, 10] - e["mtot"]) <= 1e-13 * e["mtot"] and int(hist[it, 11]) == ref.npart
        dt = min(sc.dtcourant, sc.dtforce, dterr)
    # every particle is owned by exactly one rank at the end, where the oracle has it
    seen = np.zeros(ref.npart, dtype=int)
    for o_ in outs:
        idx = o_["idx"]
        seen[idx] += 1
        dx = o_["xyzh"][:, :3] - ref.xyzh[idx, :3]
        dx -= np.round(dx)                                             # periodic unit box: the wrap may differ by a box length
        assert np.max(np.abs(dx)) < 1e-10
        assert np.max(np.abs(o_["xyzh"][:, 3] - ref.xyzh[idx, 3]) / ref.xyzh[idx, 3]) < 1e-9
    assert np.all(seen == 1)
    assert sum(int(o_["migrated"]) for o_ in outs) > 0                 # particles did change owner during the run


def test_two_gpu_shock_tube_with_boundary_particles():
    """C1 on 2 GPUs: boundary particles are active on the first density pass (deriv.f90:146) on their owner and must stay neighbour-only
    as ghosts on the other rank"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from test_halo_gloo import launch, check_against_undivided
    with tempfile.TemporaryDirectory() as d:
        launch("nccl", 2, 516, d, 29619)
        check_against_undivided(d, 2, 516)

"""Adversarial cases for the FP16 prefilter of the pair kernels (csrc/walk.cuh): the filter may pass non-neighbours but must never drop
a true pair, so the neighbour COUNTS of the density pass (q2i < R^2, dens.F90:675-679) and of the force pass (q2i < R^2 .or. q2j < R^2,
force.F90:1271-1287) have to equal the O(N^2) counts evaluated in FP64 (the reference's own check, test_neigh.f90:264-367) exactly --
also when many pairs sit within an ulp of the kernel radius, when h varies by orders of magnitude inside one target group, and when the
positions carry a large common offset."""
import numpy as np
import pytest

from phantom_b200 import setups
from oraclelib import Oracle

pytestmark = pytest.mark.gpu


def gpu(params):
    from phantom_b200.api import SphGpu
    return SphGpu(params.copy())


def counts_match(part):
    """density pass (icall 0: one pass at the given h) and force pass (at the h the density pass stored) against brute force"""
    o = Oracle(part.params)
    tot_d, _ = o.neighbour_counts_bruteforce(part, symmetric=False)
    pg = part.copy()
    g = gpu(pg.params)
    g.build_tree(pg)
    sd = g.densityiterate(pg, icall=0)
    assert sd.nactualtot == tot_d + part.npart                      # the density count includes the particle itself
    assert np.array_equal(pg.xyzh[:, :3], part.xyzh[:, :3])
    tot_f, _ = o.neighbour_counts_bruteforce(pg, symmetric=True)
    g.cons2prim_everything(pg)
    sf = g.force(pg)
    assert sf.npairs_force == tot_f
    return sd, sf


@pytest.mark.parametrize("shell2,eps", [(5, 1e-13), (5, -1e-13), (6, 3e-16), (8, -3e-16), (9, 1e-9)])
def test_lattice_pairs_on_the_kernel_edge(shell2, eps):
    # cubic lattice, periodic; 2h = sqrt(shell2) dx (1 + eps): a whole shell of neighbours sits on the edge of the kernel
    part, _ = setups.setup_test_derivs(nx=20, dissipation=False)
    dx = 1.0 / 20
    part.xyzh[:, 3] = 0.5 * np.sqrt(float(shell2)) * dx * (1.0 + eps)
    counts_match(part)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_positions_wide_h_range(seed):
    # non-periodic cloud, h log-uniform over two decades: leaf cells and target groups mix tiny and huge kernels
    rng = np.random.RandomState(seed)
    part = setups.setup_random_sphere(n=3000, gravity=False)
    n = part.npart
    part.xyzh[:, 3] = 0.02 * 10.0 ** rng.uniform(-1.0, 1.0, n)
    counts_match(part)


def test_clustered_positions_with_offset():
    # tight clumps far from the origin: the FP16 coordinates are relative to the group centre, the absolute offset must not matter
    rng = np.random.RandomState(7)
    part = setups.setup_random_sphere(n=2048, gravity=False)
    n = part.npart
    centres = rng.uniform(-1, 1, (16, 3))
    part.xyzh[:, :3] = centres[rng.randint(0, 16, n)] + 1e-3 * rng.normal(size=(n, 3)) + 1000.0
    part.xyzh[:, 3] = 1e-3 * 10.0 ** rng.uniform(-0.5, 0.5, n)
    counts_match(part)

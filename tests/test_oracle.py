"""CPU tests that PIN the oracle.

1. oracle/cpic_oracle.c (our restatement) is bit-identical to the reference's own sources
   (oracle/_ref/libcpic_ref_*.so, compiled from /root/reference) on every deck the reference
   ships, in float and double, function by function and over many steps -- including a random
   3-D state (nz > 1), which no reference test covers.
2. The reference build itself reproduces the reference's gold energy file
   (tests/energy_comparison/energies_gold.2stream-em.*, sub-sampled into tests/golden/).
3. The restatement reproduces the committed golden fixtures (tests/golden/*.npz), so the
   oracle stays pinned on the GPU box even if oracle/_ref did not travel.
"""
import os

import numpy as np
import pytest

from helpers import consts_for, random_state
from oracle.api import PARTICLE_NAMES, Consts, RefLib, Restatement

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DECKS = ["2stream-em", "custom_init", "dioctron_3d", "2particle"]

needs_ref = pytest.mark.skipif(not RefLib.available("default", "f32"), reason="oracle/_ref not built")


def same_state(a, b):
    return (all(np.array_equal(a.p[n], b.p[n]) for n in PARTICLE_NAMES) and np.array_equal(a.f, b.f)
            and np.array_equal(a.interp, b.interp) and np.array_equal(a.acc, b.acc))


@needs_ref
@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("deck", DECKS)
def test_restatement_bitwise_vs_reference_on_decks(deck, prec):
    R = RefLib(deck, prec)
    P = R.deck_params()
    k, _, _ = R.deck_consts()
    R.create_from_deck(0)
    grid = (P["nx"], P["ny"], P["nz"], P["ng"])
    s = R.get(grid=grid)
    nsteps = 120
    er = R.run(k, nsteps, energies=True)
    eo = Restatement(prec).step(s, k, 0, nsteps, energies=True)
    assert same_state(R.get(grid=grid), s)
    assert np.array_equal(er, eo)


@needs_ref
@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("grid", [(6, 5, 4), (1, 7, 3), (4, 1, 1), (3, 3, 1)])
def test_restatement_bitwise_vs_reference_3d_per_function(grid, prec):
    """Every entry point separately, on a random state with crossings and periodic wraps on
    all three axes (nz > 1 is never exercised by the reference's own tests)."""
    nx, ny, nz = grid
    s = random_state(nx, ny, nz, nppc=40, prec=prec, seed=7)
    k = consts_for(nx, ny, nz, prec)
    R = RefLib("default", prec).create(s, solver=0)
    O = Restatement(prec)
    hp = [0.5 * k.px, 0.5 * k.py, 0.5 * k.pz]
    for step in range(3):
        R.load_interpolator(); O.load_interpolator(s)
        assert np.array_equal(R.get(grid=(nx, ny, nz, 1)).interp, s.interp)
        R.clear_accumulator(); O.clear_accumulator(s)
        R.push(k); movers, crossings = O.push(s, k)
        assert movers > 0 and crossings >= movers
        r = R.get(grid=(nx, ny, nz, 1))
        assert all(np.array_equal(r.p[n], s.p[n]) for n in PARTICLE_NAMES)
        assert np.array_equal(r.acc, s.acc)
        R.unload_accumulator(k); O.unload_accumulator(s, k)
        R.advance_b(*hp); O.advance_b(s, *hp)
        R.advance_e(k.px, k.py, k.pz, k.dt_eps0); O.advance_e(s, k.px, k.py, k.pz, k.dt_eps0)
        R.advance_b(*hp); O.advance_b(s, *hp)
        assert same_state(R.get(grid=(nx, ny, nz, 1)), s)
        assert R.energies() == O.energies(s)
    # every interior cell index must still be interior (periodic wrap worked)
    c = s.p["cell"]
    ix, iy, iz = c % (nx + 2), (c // (nx + 2)) % (ny + 2), c // ((nx + 2) * (ny + 2))
    assert ix.min() >= 1 and ix.max() <= nx and iy.min() >= 1 and iy.max() <= ny and iz.min() >= 1 and iz.max() <= nz


@needs_ref
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_restatement_es1d_and_uncenter_vs_reference(prec):
    s = random_state(16, 1, 1, nppc=50, prec=prec, seed=3)
    k = consts_for(16, 1, 1, prec)
    R = RefLib("default", prec).create(s, solver=1)
    O = Restatement(prec)
    R.load_interpolator(); O.load_interpolator(s)
    R.uncenter(k.qdt_2mc); O.uncenter(s, k.qdt_2mc)
    assert all(np.array_equal(R.get(grid=(16, 1, 1, 1)).p[n], s.p[n]) for n in PARTICLE_NAMES)
    er = R.run(k, 25, energies=True)
    eo = O.step(s, k, 1, 25, energies=True)
    assert same_state(R.get(grid=(16, 1, 1, 1)), s)
    assert np.array_equal(er, eo)


@needs_ref
@pytest.mark.parametrize("prec,tol_window", [("f64", 1e-5), ("f32", 0.10)])
def test_reference_build_reproduces_gold_energy_file(prec, tol_window):
    """The reference's own regression criterion (tests/energy_comparison/2stream-em.cxx:23,45-70:
    relative error < 10 % on lines 3581..4880), and for double the much tighter 1e-5 on every
    sub-sampled line of the gold file."""
    gold = np.load(os.path.join(GOLDEN, "energies_gold_2stream-em.npz"))
    lines, g = gold["lines"], gold[prec]
    R = RefLib("2stream-em", prec)
    k, _, _ = R.deck_consts()
    R.create_from_deck(0)
    en = R.run(k, 6000, energies=True)[lines]
    rel = np.abs(en - g) / np.minimum(en, g)
    window = (lines >= 3581) & (lines < 4881)
    assert rel[window].max() < tol_window
    if prec == "f64":
        assert rel.max() < 1e-5
    else:
        assert rel[lines < 3581].max() < 0.01      # linear phase: float agrees to < 1 %


@pytest.mark.parametrize("name", ["2stream-em_f32", "2stream-em_f64", "custom_init_f32", "dioctron_3d_f32",
                                  "2particle_f32", "random3d_f32", "random3d_f64"])
def test_restatement_reproduces_golden_fixtures(name):
    """Fixtures were generated from the REFERENCE build by tests/golden/make_golden.py."""
    from oracle.api import State
    z = np.load(os.path.join(GOLDEN, f"state_{name}.npz"))
    prec = name.split("_")[-1]
    nx, ny, nz, ng, nsteps, solver = [int(v) for v in z["meta"]]
    s = State(nx, ny, nz, ng, len(z["p0_cell"]), prec)
    for n in PARTICLE_NAMES:
        s.p[n][:] = z["p0_" + n]
    s.f[:] = z["f0"]
    k = Consts(**{n: float(v) for n, v in zip("qdt_2mc cdt_dx cdt_dy cdt_dz qsp dx dy dz dt px py pz dt_eps0".split(),
                                              z["consts"])})
    en = Restatement(prec).step(s, k, solver, nsteps, energies=True)
    for n in PARTICLE_NAMES:
        assert np.array_equal(s.p[n], z["p1_" + n]), n
    assert np.array_equal(s.f, z["f1"])
    assert np.array_equal(en, z["energies"])


def test_float_history_summation_order_sensitivity():
    """Justifies the float tolerance used for the GPU (tests/test_gpu_parity.py): merely
    permuting the particle order in the SERIAL oracle -- the only change is the order in which
    the float accumulator sums the same deposits -- moves the early E-energy minima of the
    2stream-em history by > 0.3 % relative, while the history stays within 0.5 % of the local
    oscillation envelope before saturation.  Inside the reference's comparison window (chaotic,
    post-saturation) this permutation lands at 13.7 % -- i.e. the reference's own 10 % float
    criterion does not survive a reordering of its own sums (its CI only runs double) -- so
    float histories from a different summation order are held to 25 % there."""
    from oracle.api import State
    z = np.load(os.path.join(GOLDEN, "state_2stream-em_f32.npz"))
    gold = np.load(os.path.join(GOLDEN, "energies_gold_2stream-em.npz"))
    lines, g, env = gold["lines"], gold["f32"], gold["env_f32"]
    nx, ny, nz, ng, _, solver = [int(v) for v in z["meta"]]
    k = Consts(**{n: float(v) for n, v in zip("qdt_2mc cdt_dx cdt_dy cdt_dz qsp dx dy dz dt px py pz dt_eps0".split(),
                                              z["consts"])})
    perm = np.random.default_rng(3).permutation(len(z["p0_cell"]))
    s = State(nx, ny, nz, ng, len(perm), "f32")
    for n in PARTICLE_NAMES:
        s.p[n][:] = z["p0_" + n][perm]
    s.f[:] = z["f0"]
    en = Restatement("f32").step(s, k, solver, 6000, energies=True)[lines]
    rel = np.abs(en - g) / np.minimum(en, g)
    erel = np.abs(en - g) / env
    window = (lines >= 3581) & (lines < 4881)
    assert rel[lines < 3581].max() > 3e-3          # order alone already breaks a tight min-relative bound
    assert erel[lines < 3581].max() < 5e-3
    assert 0.10 < rel[window].max() < 0.25

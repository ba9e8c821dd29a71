"""Generate the committed golden fixtures from the REFERENCE build (oracle/_ref).

Run in the build container (needs /root/reference for the gold energy files and
oracle/_ref built by `make -C oracle ref`):   python tests/golden/make_golden.py

Outputs (small, committed):
  energies_gold_2stream-em.npz   every 25th line of the reference's own gold files
                                 tests/energy_comparison/energies_gold.2stream-em.{float,double}
                                 (columns E, B) plus the whole comparison window stride 10
  state_<deck>_<prec>.npz        initial state, step constants, state after N steps and the
                                 per-step energies, all produced by the reference's sources
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from helpers import consts_for, random_state  # noqa: E402
from oracle.api import CONST_NAMES, PARTICLE_NAMES, RefLib  # noqa: E402

REF = "/root/reference/tests/energy_comparison/energies_gold.2stream-em"


def gold_energies():
    lines = np.unique(np.concatenate([np.arange(0, 6000, 25), np.arange(3581, 4881, 10), [5999]]))
    out = {"lines": lines}
    for prec, suffix in (("f32", "float"), ("f64", "double")):
        g = np.loadtxt(f"{REF}.{suffix}")
        out[prec] = g[lines, 2:4]
        # local oscillation envelope: running maximum over +-64 lines.  The E energy of this deck
        # oscillates between ~2e-11 and ~2e-16 in the linear phase; at the minima the value is pure
        # summation-order noise, so differences are judged against the envelope, not the minimum.
        from scipy.ndimage import maximum_filter1d
        out["env_" + prec] = np.stack([maximum_filter1d(g[:, c], 129) for c in (2, 3)], axis=1)[lines]
    np.savez_compressed(os.path.join(HERE, "energies_gold_2stream-em.npz"), **out)


def deck_state(deck, prec, nsteps, solver=0):
    R = RefLib(deck, prec)
    P = R.deck_params()
    k, _, _ = R.deck_consts()
    R.create_from_deck(solver)
    grid = (P["nx"], P["ny"], P["nz"], P["ng"])
    save(f"{deck}_{prec}", R, k, grid, nsteps, solver)


def random_state_fixture(prec):
    nx, ny, nz = 6, 5, 4
    s = random_state(nx, ny, nz, nppc=24, prec=prec, seed=11)
    k = consts_for(nx, ny, nz, prec)
    R = RefLib("default", prec).create(s, solver=0)
    save(f"random3d_{prec}", R, k, (nx, ny, nz, 1), 8, 0)


def save(name, R, k, grid, nsteps, solver):
    s0 = R.get(grid=grid)
    en = R.run(k, nsteps, energies=True)
    s1 = R.get(grid=grid)
    d = {"meta": np.array(list(grid) + [nsteps, solver]), "consts": np.array([getattr(k, n) for n in CONST_NAMES]),
         "f0": s0.f, "f1": s1.f, "energies": en, "acc1": s1.acc, "interp1": s1.interp}
    for n in PARTICLE_NAMES:
        d["p0_" + n] = s0.p[n]
        d["p1_" + n] = s1.p[n]
    np.savez_compressed(os.path.join(HERE, f"state_{name}.npz"), **d)
    print(name, "particles", len(s0.p["cell"]), "steps", nsteps)


if __name__ == "__main__":
    gold_energies()
    deck_state("2stream-em", "f32", 40)
    deck_state("2stream-em", "f64", 40)
    deck_state("custom_init", "f32", 30)
    deck_state("dioctron_3d", "f32", 10)
    deck_state("2particle", "f32", 200)
    random_state_fixture("f32")
    random_state_fixture("f64")

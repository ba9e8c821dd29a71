"""Developer check of the experimental PUSH2_PLACE build (placement by the new cell): parity against the oracle on a
grid above the block-private limit with store headroom, then timing.  CPIC_LIB must point at the place build."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cabanapic_b200 as cp  # noqa: E402
from cabanapic_b200 import decks  # noqa: E402
from helpers import canonical_order, consts_for, random_state  # noqa: E402
from oracle.api import PARTICLE_NAMES, Restatement  # noqa: E402


def parity():
    nx, ny, nz = 17, 9, 5
    for seed, nppc in ((8, 37), (9, 64)):
        s = random_state(nx, ny, nz, nppc=nppc, prec="f32", seed=seed)
        k = consts_for(nx, ny, nz, "f32")
        kk = cp.Consts(**k.to_dict())
        O = Restatement("f32")
        # (a) teacher-forced pushes, segment form kept between them (fields never advance: same interpolators)
        with cp.Context(nx, ny, nz, 1, max_particles=int(s.np * 1.6) + 8192, real=np.float32) as c:
            c.upload_particles(s.p); c.upload_fields(s.f)
            for it in range(4):
                c.load_interpolator_array(); c.clear_accumulator_array(); c.push_reorder(kk)
                O.load_interpolator(s); O.clear_accumulator(s); O.push(s, k)
                acc = c.download_accumulators()
                scale = np.abs(s.acc).max()
                assert np.abs(acc - s.acc).max() <= 2e-5 * scale, ("acc", it)
            assert c.num_particles == s.np
            p = c.download_particles()
            a, b = canonical_order(p), canonical_order(s.p)
            for n in PARTICLE_NAMES:
                assert np.array_equal(p[n][a], s.p[n][b]), (n, "teacher-forced")
            assert np.all(np.diff(p["cell"]) >= 0)
        # (b) 7 whole steps in one call against the oracle
        s = random_state(nx, ny, nz, nppc=nppc, prec="f32", seed=seed)
        with cp.Context(nx, ny, nz, 1, max_particles=int(s.np * 1.6) + 8192, real=np.float32) as c:
            c.upload_particles(s.p); c.upload_fields(s.f)
            c.step(kk, 7, cp.SORT_FUSED, False)
            p, f = c.download_particles(), c.download_fields()
        O.step(s, k, 0, 7)
        assert len(p["cell"]) == s.np
        a, b = canonical_order(p), canonical_order(s.p)
        assert np.mean(p["cell"][a] == s.p["cell"][b]) > 0.999
        same = p["cell"][a] == s.p["cell"][b]
        for n in ("dx", "dy", "dz", "ux", "uy", "uz"):
            assert np.abs(p[n][a][same] - s.p[n][b][same]).max() < 5e-5, n
        scale = np.abs(s.f).max(axis=1, keepdims=True) + 1e-30
        assert (np.abs(f - s.f) / scale).max() < 2e-4
        print(f"parity ok (seed {seed}, nppc {nppc})", flush=True)


def timing(nx=256, ny=256, nz=64, nppc=64, steps=5):
    d = decks.uniform_plasma(nx, ny, nz, nppc)
    k, _, we = d.consts()
    n = d.num_particles
    c = cp.Context(nx, ny, nz, 1, max_particles=int(n * 1.15), real=np.float32)
    c.init_uniform_plasma(0, n, nx, ny, nz, nppc, weight=we)
    c.upload_fields(d.initial_fields())
    for s in range(steps):
        c.step(k, 1, cp.SORT_FUSED, False); c.sync()
        print(f"placing push, step {s}: {c.last_ms(0):8.3f} ms  {56 * n / c.last_ms(0) / 1e6:8.1f} GB/s", flush=True)
    assert c.num_particles == n
    c.close()


if __name__ == "__main__":
    parity()
    timing()

"""Developer probe (not part of the product): per-step device time of the reordering push
(cpic_step with CPIC_SORT_FUSED) next to the in-place push one step after a sort.
Usage: python tools/probe_reorder.py [nx ny nz nppc [steps]]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cabanapic_b200 as cp  # noqa: E402
from cabanapic_b200 import decks  # noqa: E402


def main():
    a = sys.argv[1:]
    nx, ny, nz, nppc = [int(v) for v in (a[:4] if len(a) >= 4 else (256, 256, 32, 64))]
    steps = int(a[4]) if len(a) > 4 else 5
    mode = a[5] if len(a) > 5 else "both"
    d = decks.uniform_plasma(nx, ny, nz, nppc)
    k, _, we = d.consts()
    n = d.num_particles
    c = cp.Context(nx, ny, nz, 1, max_particles=n, real=np.float32)
    c.init_uniform_plasma(0, n, nx, ny, nz, nppc, weight=we)
    c.upload_fields(d.initial_fields())
    if os.environ.get("CPIC_FP", "strict") == "contract":
        c.set_modes(cp.FP_CONTRACT, 3)
    if mode in ("both", "inplace"):
        for s in range(steps):
            c.sort_particles()
            c.step(k, 1, 0, False)          # one step of drift
            c.load_interpolator_array(); c.clear_accumulator_array(); c.push(k); c.sync()
            print(f"in-place push, one step after a sort: {c.last_ms(0):8.3f} ms  {56 * n / c.last_ms(0) / 1e6:8.1f} GB/s   (that sort: {c.last_ms(1):8.3f} ms)", flush=True)
        c.sort_particles()
        c.load_interpolator_array(); c.clear_accumulator_array(); c.push(k); c.sync()
        print(f"in-place push, freshly sorted: {c.last_ms(0):8.3f} ms  {56 * n / c.last_ms(0) / 1e6:8.1f} GB/s", flush=True)
    if mode in ("both", "reorder"):
        for s in range(steps):
            if s == steps - 1 and os.environ.get("CPIC_KO_LAST"):      # developer knock-outs (a -DPUSH3_KO=1 build), last step only
                os.environ["CPIC_KO"] = os.environ["CPIC_KO_LAST"]
            c.step(k, 1, cp.SORT_FUSED, False); c.sync()
            print(f"reordering push, step {s}: {c.last_ms(0):8.3f} ms  {56 * n / c.last_ms(0) / 1e6:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()

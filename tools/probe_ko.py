"""Developer probe (not part of the product): steady-state knock-out timing of the reordering push.
Needs a library built with -DPUSH2_KO_RT (make var NAME=kort DEFS=-DPUSH2_KO_RT; CPIC_LIB=...).  For every mask:
fresh plasma, `warm` normal reordering steps, then ONE step with the run-time knock-out mask, timed.
Usage: python tools/probe_ko.py nx ny nz nppc warm mask [mask ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cabanapic_b200 as cp  # noqa: E402
from cabanapic_b200 import decks  # noqa: E402


def main():
    a = sys.argv[1:]
    nx, ny, nz, nppc, warm = [int(v) for v in a[:5]]
    masks = [int(v) for v in a[5:]] or [0]
    d = decks.uniform_plasma(nx, ny, nz, nppc)
    k, _, we = d.consts()
    n = d.num_particles
    c = cp.Context(nx, ny, nz, 1, max_particles=n, real=np.float32)
    f0 = d.initial_fields()
    for m in masks:
        os.environ["CPIC_PUSH2_KO"] = "0"
        c.init_uniform_plasma(0, n, nx, ny, nz, nppc, weight=we)
        c.upload_fields(f0)
        c.step(k, warm, cp.SORT_FUSED, False)
        c.sync()
        base = c.last_ms(0)
        os.environ["CPIC_PUSH2_KO"] = str(m)
        c.step(k, 1, cp.SORT_FUSED, False)
        c.sync()
        t = c.last_ms(0)
        print(f"ko {m:4d}: {t:8.3f} ms  ({1e6 * t / n:6.2f} ps/particle; the normal step before it {base:8.3f} ms)", flush=True)
    os.environ["CPIC_PUSH2_KO"] = "0"


if __name__ == "__main__":
    main()

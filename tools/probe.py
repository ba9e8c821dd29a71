"""Developer probe (not part of the product): device timings of the push kernel under the
different deposit / floating-point modes, of the sort and of the field side, on the synthetic
uniform plasma.  Usage: python tools/probe.py [nx ny nz nppc]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cabanapic_b200 as cp  # noqa: E402
from cabanapic_b200 import decks  # noqa: E402


def main():
    nx, ny, nz, nppc = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (128, 128, 128, 64))]
    d = decks.uniform_plasma(nx, ny, nz, nppc)
    k, _, we = d.consts()
    n = d.num_particles
    c = cp.Context(nx, ny, nz, 1, max_particles=n, real=np.float32)
    c.init_uniform_plasma(0, n, nx, ny, nz, nppc, weight=we)
    c.upload_fields(d.initial_fields())
    out = {"grid": [nx, ny, nz], "nppc": nppc, "np": n}

    def timed_push(reps=3):
        ts = []
        for _ in range(reps):
            c.load_interpolator_array(); c.clear_accumulator_array()
            c.push(k)
            c.sync()
            ts.append(c.last_ms(0))
        return min(ts)

    for fp, fpn in ((cp.FP_STRICT, "strict"), (cp.FP_CONTRACT, "contract")):
        for dep, depn in ((1, "atomic"), (2, "atomic_v4"), (3, "warp")):
            c.set_modes(fp, dep)
            c.sort_particles()
            ms = timed_push()
            out[f"push_sorted_{fpn}_{depn}_ms"] = ms
            print(f"push sorted   {fpn:8s} {depn:9s}: {ms:8.3f} ms  {n / ms / 1e6:9.1f} Gp/s-ish(M/ms)  "
                  f"{56 * n / ms / 1e6:8.1f} GB/s", flush=True)
    c.set_modes(cp.FP_STRICT, 3)
    # drift without sorting: how quickly does the warp-uniform fast path decay?
    c.sort_particles()
    for s in range(24):
        c.load_interpolator_array(); c.clear_accumulator_array(); c.push(k); c.sync()
        print(f"push step {s} since sort: {c.last_ms(0):8.3f} ms", flush=True)
        out[f"push_since_sort_{s}_ms"] = c.last_ms(0)
    c.sort_particles(); c.sync()
    print(f"sort (drifted 24 steps): {c.last_ms(1):8.3f} ms")
    out["sort_drift24_ms"] = c.last_ms(1)
    c.sort_particles(); c.sync()
    print(f"sort (already sorted) : {c.last_ms(1):8.3f} ms")
    out["sort_sorted_ms"] = c.last_ms(1)
    # whole fused steps, sort every step
    for si in (1, 4, 8, 16, 0):
        c.sort_particles()
        c.step(k, 2, si, False); c.sync()
        c.step(k, 32, si, False); c.sync()
        ms = c.last_ms(3) / 32
        print(f"fused step sort_interval={si}: {ms:8.3f} ms/step  {n / ms / 1e6:8.1f} Mp/ms  {56 * n / ms / 1e6:8.1f} GB/s")
        out[f"step_sort{si}_ms"] = ms
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()

"""Developer sweep: push-kernel time for library variants (PUSH_MIN_BLOCKS) x prefetch x fp mode.
Each configuration runs in a fresh process because the knobs are read at context creation."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json, numpy as np
sys.path.insert(0, %r)
import cabanapic_b200 as cp
from cabanapic_b200 import decks
nx, ny, nz, nppc = 128, 128, 128, 64
d = decks.uniform_plasma(nx, ny, nz, nppc); k, _, we = d.consts(); n = d.num_particles
c = cp.Context(nx, ny, nz, 1, max_particles=n, real=np.float32)
c.init_uniform_plasma(0, n, nx, ny, nz, nppc, weight=we); c.upload_fields(d.initial_fields())
res = {}
for fp, name in ((cp.FP_STRICT, "strict"), (cp.FP_CONTRACT, "contract")):
    c.set_modes(fp, 3); c.sort_particles(); ts = []
    for _ in range(4):
        c.load_interpolator_array(); c.clear_accumulator_array(); c.push(k); c.sync(); ts.append(c.last_ms(0))
    res[name] = min(ts)
print(json.dumps(res))
''' % ROOT


def main():
    for th in (4, 6, 10, 16, 33):
        for rounds in (1, 2, 4):
            env = dict(os.environ, CPIC_DEP_THRESH=str(th), CPIC_DEP_ROUNDS=str(rounds))
            r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
            out = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
            print(f"thresh={th} rounds={rounds}: {out}", flush=True)


if __name__ == "__main__":
    main()

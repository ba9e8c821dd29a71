"""Generate the long energy histories of the reference build (oracle/_ref) for BASELINE configs[2] (C3, the dioctron deck:
decks/dioctron_3d.cxx:167-200, 20 000 steps) -- the E / B field energies after every 50th step (steps 50, 100, ...: what the driver's energies.txt holds with
CPIC_ENERGY_INTERVAL=50) the reference's own sources
produce in float.  Run in the build container after `make -C oracle ref`:   python tests/golden/make_long_histories.py
Output: tests/golden/history_dioctron_3d_f32.npz (committed, small)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.api import RefLib  # noqa: E402


def main():
    R = RefLib("dioctron_3d", "f32")
    P = R.deck_params()
    k, _, _ = R.deck_consts()
    R.create_from_deck(0)
    n = P["num_steps"]
    t0 = time.time()
    en = R.run(k, n, energies=True)
    stride = 50
    np.savez_compressed(os.path.join(HERE, "history_dioctron_3d_f32.npz"), steps=np.arange(stride, n + 1, stride), energies=en[stride - 1::stride],
                        num_steps=n, stride=stride)
    print(f"dioctron_3d f32: {n} steps in {time.time() - t0:.1f} s; E {en[0, 0]:.6e} -> {en[-1, 0]:.6e}, B {en[0, 1]:.6e} -> {en[-1, 1]:.6e}")


if __name__ == "__main__":
    main()

"""Developer probe: full 6000-step field-energy history of the reference's 2stream-em regression
deck (tests/energy_comparison/2stream-em.cxx) on the GPU, per precision and deposit mode, saved to
gpurun_out/energy_history.npz for offline comparison with the gold files."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cabanapic_b200 as cp  # noqa: E402
from cabanapic_b200 import decks  # noqa: E402

out = {}
for real, pn in ((np.float32, "f32"), (np.float64, "f64")):
    for dep in (1, 2, 3):
        for fp, fpn in ((cp.FP_STRICT, "strict"), (cp.FP_CONTRACT, "contract")):
            for sort in (0, 1):
                if pn == "f64" and (dep != 3 or sort):
                    continue
                sim = cp.Simulation(decks.two_stream_em(real), fp_mode=fp, deposit_mode=dep)
                en = sim.run(6000, sort_interval=sort, energies=True)
                sim.close()
                out[f"{pn}_dep{dep}_{fpn}_sort{sort}"] = en
                print(pn, dep, fpn, sort, en[0], en[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/energy_history.npz", **out)

"""Known-answer fixture from the reference's own data: tests/manual_tests/test/2-particle/partloc (200 001 records
`t x1 v1 x2 v2` written by example/example.cpp:131-132,276-277 through dump_particles, src/helpers.h:26-63, for
decks/2particle.cxx in float).  Keeps the first 3001 records, every 20th plus the last one.
Run in the build container (needs /root/reference):   python tests/golden/make_partloc_fixture.py"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/tests/manual_tests/test/2-particle/partloc"

rows = []
with open(SRC) as fh:
    for line in fh:
        if line.startswith("#"):
            continue
        rows.append([float(v) for v in line.split()])
        if len(rows) == 3001:
            break
a = np.array(rows)
steps = np.unique(np.concatenate([np.arange(0, 3001, 20), [3000]]))
np.savez_compressed(os.path.join(HERE, "partloc_2particle.npz"), steps=steps, records=a[steps],
                    source="reference tests/manual_tests/test/2-particle/partloc, records 0..3000")
print("wrote", len(steps), "records; x1 range", a[:, 1].min(), a[:, 1].max())

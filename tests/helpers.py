"""Shared helpers for the test-suite (random states, constants, canonical particle order)."""
import numpy as np

# ------------------------------------------------------------------------- shared helpers
PREC = {"f32": np.float32, "f64": np.float64}


def random_state(nx, ny, nz, nppc, prec="f32", seed=0, uth=0.3, field_amp=0.5):
    """A random but physical state: particles uniformly spread over the interior cells with
    thermal momenta large enough that a good fraction crosses cells (and wraps periodically)
    in every direction; smooth-ish random E and cB with periodic ghosts."""
    from oracle.api import State
    rng = np.random.default_rng(seed)
    R = PREC[prec]
    npart = nx * ny * nz * nppc
    s = State(nx, ny, nz, 1, npart, prec)
    c = np.arange(npart) // nppc
    ix, iy, iz = c % nx, (c // nx) % ny, c // (nx * ny)
    s.p["cell"][:] = (ix + 1) + (nx + 2) * ((iy + 1) + (ny + 2) * (iz + 1))
    for n in ("dx", "dy", "dz"):
        s.p[n][:] = rng.uniform(-1, 1, npart).astype(R)
    for n in ("ux", "uy", "uz"):
        s.p[n][:] = (uth * rng.standard_normal(npart)).astype(R)
    s.p["w"][:] = R(0.01) * (1 + rng.uniform(0, 1, npart)).astype(R)
    f = (field_amp * rng.standard_normal((9, nz + 2, ny + 2, nx + 2))).astype(R)
    f[6:] = 0
    # periodic ghosts
    for a in range(6):
        v = f[a]
        v[:, :, 0] = v[:, :, nx]; v[:, :, nx + 1] = v[:, :, 1]
        v[:, 0, :] = v[:, ny, :]; v[:, ny + 1, :] = v[:, 1, :]
        v[0, :, :] = v[nz, :, :]; v[nz + 1, :, :] = v[1, :, :]
    s.f[:] = f.reshape(9, -1)
    return s


def consts_for(nx, ny, nz, prec="f32", cfl=0.7, qdt_2mc=-0.05):
    """Plausible step constants for a random-state test (any values work for parity)."""
    from oracle.api import Consts
    R = PREC[prec]
    dx = dy = dz = R(0.1)
    dt = R(cfl * 0.1 / np.sqrt(3.0))
    cdt = R(dt / dx)
    return Consts(qdt_2mc=float(R(qdt_2mc)), cdt_dx=float(cdt), cdt_dy=float(cdt), cdt_dz=float(cdt), qsp=-1.0,
                  dx=float(dx), dy=float(dy), dz=float(dz), dt=float(dt),
                  px=float(cdt) if nx > 1 else 0.0, py=float(cdt) if ny > 1 else 0.0,
                  pz=float(cdt) if nz > 1 else 0.0, dt_eps0=float(dt))


def canonical_order(p):
    """Order-independent view of a particle set (after a sort the order is arbitrary)."""
    return np.lexsort((p["uz"], p["uy"], p["ux"], p["dz"], p["dy"], p["dx"], p["cell"]))

"""Host-side mirror of the reference's deck interface for tests and bench.py.

Reference: ``class _Input_Deck`` (src/input/deck.h:161-348), the driver-side constants
(example/example.cpp:77-113,179-181) and the particle initialisers of
src/input/deck.h:111-153 (default, y-oriented) and decks/custom_init.cxx:29-106
(x-oriented).  Decks written in C++ (decks/*.cxx) drop in through the C++ facade in
``include/cabanapic/``; this module exists so Python tests and the benchmark can set up
the same problems without a C++ compile.  All arithmetic mirrors the reference's mixed
real_t/double promotions so the numbers are bit-identical (checked in tests/test_decks.py
against the reference build).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

from ._lib import CONST_NAMES, Consts


@dataclasses.dataclass
class Deck:
    """Field names and defaults follow _Input_Deck (src/input/deck.h:254-309)."""
    real: type = np.float32
    de: float = 1.0
    ec: float = 1.0
    me: float = 1.0
    mu: float = 1.0
    c: float = 1.0
    eps: float = 1.0
    qsp: float = -1.0
    n0: float = 1.0
    nx: int = 16
    ny: int = 1
    nz: int = 1
    num_ghosts: int = 1
    nppc: int = 1
    dt: float = 1.0
    num_steps: int = 2
    len_x_global: float = 1.0
    len_y_global: float = 1.0
    len_z_global: float = 1.0
    Npe: float = -1.0
    Ne: float = -1.0
    v0: float = 1.0
    num_particles: int = -1
    perform_uncenter: bool = False
    init: str = "default"      # which particle initialiser: default | custom_init | uniform
    name: str = "deck"
    # derived (derive_params)
    dx: float = 0.0
    dy: float = 0.0
    dz: float = 0.0
    num_cells: int = 0

    def R(self, v):
        return self.real(v)

    @staticmethod
    def courant_length(real, lx, ly, lz, nx, ny, nz):
        """src/input/deck.h:191-198 (axes with one cell are skipped)."""
        w1 = real(0)
        if nx > 1:
            w0 = real(nx) / real(lx); w1 = real(w1 + w0 * w0)
        if ny > 1:
            w0 = real(ny) / real(ly); w1 = real(w1 + w0 * w0)
        if nz > 1:
            w0 = real(nz) / real(lz); w1 = real(w1 + w0 * w0)
        return real(np.sqrt(real(real(1) / w1)))

    def derive_params(self):
        """src/input/deck.h:323-348."""
        R = self.real
        self.dx = R(R(self.len_x_global) / R(self.nx))
        self.dy = R(R(self.len_y_global) / R(self.ny))
        self.dz = R(R(self.len_z_global) / R(self.nz))
        g = 2 * self.num_ghosts
        self.num_cells = (self.nx + g) * (self.ny + g) * (self.nz + g)
        if self.num_particles < 0:
            self.num_particles = self.nx * self.ny * self.nz * self.nppc
            if self.Ne < 0:
                self.Ne = R(self.num_particles)
        if self.Npe < 0:
            self.Npe = R(R(R(R(self.n0) * R(self.len_x_global)) * R(self.len_y_global)) * R(self.len_z_global))
        return self

    def consts(self):
        """Step constants in real_t, as example/example.cpp:77-113,179-181 computes them.
        Returns (Consts, dxp, we)."""
        R = self.real
        self.derive_params()
        dxp = R(np.float32(2.0) / np.float32(self.nppc))
        dx, dy, dz = R(self.dx), R(self.dy), R(self.dz)
        dt, c, eps0 = R(self.dt), R(self.c), R(self.eps)
        Npe = R(self.Npe)
        Ne = int(R(self.Ne))              # size_t Ne = deck.Ne
        qsp, me = R(self.qsp), R(self.me)
        qdt_2mc = R(R(qsp * dt) / R(R(R(2) * me) * c))
        cdt_dx = R(R(c * dt) / dx)
        cdt_dy = R(R(c * dt) / dy)
        cdt_dz = R(R(c * dt) / dz)
        dt_eps0 = R(dt / eps0)
        frac = R(1.0)
        we = R(Npe / R(Ne))
        px = R(R(R(frac * c) * dt) / dx) if self.nx > 1 else R(0)
        py = R(R(R(frac * c) * dt) / dy) if self.ny > 1 else R(0)
        pz = R(R(R(frac * c) * dt) / dz) if self.nz > 1 else R(0)
        vals = dict(qdt_2mc=qdt_2mc, cdt_dx=cdt_dx, cdt_dy=cdt_dy, cdt_dz=cdt_dz, qsp=qsp, dx=dx, dy=dy, dz=dz,
                    dt=dt, px=px, py=py, pz=pz, dt_eps0=dt_eps0)
        return Consts(**{n: float(vals[n]) for n in CONST_NAMES}), float(dxp), float(we)

    # ------------------------------------------------------------------ initial conditions
    def initial_particles(self):
        """Run this deck's particle initialiser on the host; returns the 8 member arrays."""
        _, dxp, we = self.consts()
        if self.init == "default":
            return _init_two_stream(self, dxp, we, x_oriented=False)
        if self.init == "custom_init":
            return _init_two_stream(self, dxp, we, x_oriented=True)
        if self.init == "uniform":
            return uniform_plasma_particles(self, we)
        raise ValueError(f"unknown initialiser {self.init!r}")

    def initial_fields(self):
        """Default Field_Initializer: everything zero (src/input/deck.h:54-65)."""
        self.derive_params()
        return np.zeros((9, self.num_cells), dtype=self.real)


def _init_two_stream(d: Deck, dxp, we, x_oriented):
    """Counter-streaming beams with a 1e-4 sinusoidal ux perturbation.

    default (y line):  src/input/deck.h:111-153
    custom_init (x line): decks/custom_init.cxx:66-100
    Intermediates are double, as in the reference (the literals are doubles); ``x`` is
    rounded to real_ first because it is declared ``real_ x``.
    """
    R = d.real
    n = d.num_particles
    nppc, nx, ny, ng = d.nppc, d.nx, d.ny, d.num_ghosts
    pi2 = np.arange(n, dtype=np.int64)
    pi = pi2 // 2
    sign = np.where(pi2 % 2 == 0, 1, -1).astype(np.int64)
    pic = (2 * pi) % nppc
    dxp_r = R(dxp)
    # real_ x = pic*dxp + 0.5*dxp - 1.0  : pic*dxp in real_ (size_t/int -> real_), rest in double
    x = (pic.astype(R) * dxp_r).astype(np.float64) + 0.5 * np.float64(dxp_r) - 1.0
    x = x.astype(R)
    pre_ghost = (2 * pi) // nppc
    v0 = R(d.v0)
    gam = R(1.0 / math.sqrt(1.0 - float(R(v0 * v0))))
    p = {}
    zeros = np.zeros(n, dtype=R)
    if not x_oriented:
        p["dx"], p["dy"], p["dz"] = zeros.copy(), x, zeros.copy()
        cell = pre_ghost * (nx + 2) + (nx + 2) * (ny + 2) + (nx + 2) + 1
        arg = (x.astype(np.float64) + 1.0 + (pre_ghost * 2).astype(np.float64)) / np.float64(2 * ny)
        na = (0.0001 * np.sin(2.0 * 3.1415926 * arg)).astype(R)
        # sign*v0*gam*(1.0+na*sign): (sign*v0)*gam in real_, (1.0 + na*sign) in double
        lead = ((sign.astype(R) * v0).astype(R) * gam).astype(R)
        ux = lead.astype(np.float64) * (1.0 + (na * sign.astype(R)).astype(R).astype(np.float64))
    else:
        p["dx"], p["dy"], p["dz"] = x, zeros.copy(), zeros.copy()
        ix = pre_ghost + 1
        cell = ix + (nx + 2 * ng) * (1 + (ny + 2 * ng) * 1)
        arg = (x.astype(np.float64) + 1.0 + (ix * 2).astype(np.float64)) / np.float64(2 * nx)
        nax = (0.0001 * np.sin(2.0 * 3.1415926 * arg)).astype(R)
        lead = ((sign.astype(R) * v0).astype(R) * gam).astype(R)
        ux = lead.astype(np.float64) * (1.0 + nax.astype(np.float64))
    p["ux"] = ux.astype(R)
    p["uy"], p["uz"] = zeros.copy(), zeros.copy()
    p["w"] = np.full(n, R(we), dtype=R)
    p["cell"] = cell.astype(np.int32)
    return p


# --------------------------------------------------------------------------------- Philox
def philox4x32(counter_lo: np.ndarray, key: int, stream: int = 0, rounds: int = 10):
    """Philox-4x32-10 counter-based RNG (Salmon et al., SC'11), vectorised over counters.

    counter = (counter_lo & 0xffffffff, counter_lo >> 32, stream, 0), key = (key lo, key hi).
    Returns four uint32 arrays.  Used so that every rank / the CPU oracle / the GPU all see
    the same synthetic particles for a given (seed, particle index).
    """
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = 0x9E3779B9, 0xBB67AE85
    c = counter_lo.astype(np.uint64)
    x0 = (c & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    x1 = (c >> np.uint64(32)).astype(np.uint64)
    x2 = np.full_like(x0, np.uint64(stream & 0xFFFFFFFF))
    x3 = np.zeros_like(x0)
    k0, k1 = key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(rounds):
        p0 = M0 * x0
        p1 = M1 * x2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        x0, x1, x2, x3 = (hi1 ^ x1 ^ np.uint64(k0)) & mask, lo1, (hi0 ^ x3 ^ np.uint64(k1)) & mask, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return x0.astype(np.uint32), x1.astype(np.uint32), x2.astype(np.uint32), x3.astype(np.uint32)


def _u01(u32):
    """uint32 -> float64 uniform in (0,1)."""
    return (u32.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)


def uniform_plasma_chunk(d: Deck, we, first: int, count: int, seed: int = 12345, vth=(0.1, 0.1, 0.1)):
    """Particles [first, first+count) of the synthetic uniform thermal plasma.

    Particle k sits in interior cell k // nppc (x fastest, as the reference's initialisers lay
    them out, i.e. the store starts cell-sorted); offsets ~ U(-1,1); momenta ~ N(0, vth) via
    Box-Muller; all from Philox(seed, k), so any rank can generate any slice independently.
    """
    R = d.real
    k = np.arange(first, first + count, dtype=np.int64)
    a = philox4x32(k, seed, stream=0)
    b = philox4x32(k, seed, stream=1)
    p = {}
    p["dx"] = (2.0 * _u01(a[0]) - 1.0).astype(R)
    p["dy"] = (2.0 * _u01(a[1]) - 1.0).astype(R)
    p["dz"] = (2.0 * _u01(a[2]) - 1.0).astype(R)
    r1 = np.sqrt(-2.0 * np.log(_u01(a[3])))
    r2 = np.sqrt(-2.0 * np.log(_u01(b[0])))
    t1 = 2.0 * np.pi * _u01(b[1])
    t2 = 2.0 * np.pi * _u01(b[2])
    p["ux"] = (vth[0] * r1 * np.cos(t1)).astype(R)
    p["uy"] = (vth[1] * r1 * np.sin(t1)).astype(R)
    p["uz"] = (vth[2] * r2 * np.cos(t2)).astype(R)
    p["w"] = np.full(count, R(we), dtype=R)
    c = k // d.nppc
    ix = c % d.nx
    iy = (c // d.nx) % d.ny
    iz = c // (d.nx * d.ny)
    g = d.num_ghosts
    p["cell"] = ((ix + g) + (d.nx + 2 * g) * ((iy + g) + (d.ny + 2 * g) * (iz + g))).astype(np.int32)
    return p


def uniform_plasma_particles(d: Deck, we, seed: int = 12345, vth=(0.1, 0.1, 0.1)):
    return uniform_plasma_chunk(d, we, 0, d.num_particles, seed, vth)


# --------------------------------------------------------------------------------- decks
def two_stream_em(real=np.float32) -> Deck:
    """tests/energy_comparison/2stream-em.cxx:80-114 (the gold-file deck; == the built-in
    default deck, src/input/deck.h:431-466): 1x32x1, nppc 100, 6000 steps, default initialiser."""
    R = real
    d = Deck(real=real, name="2stream-em", nx=1, ny=32, nz=1, num_steps=6000, nppc=100, init="default")
    d.v0 = R(0.0866025403784439)
    gam = R(1.0 / math.sqrt(1.0 - float(R(d.v0 * d.v0))))
    d.len_x_global = R(1.0)
    d.len_y_global = R(0.628318530717959 * _gam_sqrt_gam(gam))
    d.len_z_global = R(1.0)
    d.dt = R(0.99 * float(Deck.courant_length(R, d.len_x_global, d.len_y_global, d.len_z_global, d.nx, d.ny, d.nz))
             / float(R(d.c)))
    d.n0 = R(2.0)
    return d


def custom_init(real=np.float32) -> Deck:
    """decks/custom_init.cxx:108-142: 32x1x1, nppc 100, 30 steps, x-oriented initialiser.
    Npe is evaluated before n0 is set to 2 (decks/custom_init.cxx:133,141)."""
    R = real
    d = Deck(real=real, name="custom_init", nx=32, ny=1, nz=1, num_steps=30, nppc=100, init="custom_init")
    d.v0 = R(0.0866025403784439)
    gam = R(1.0 / math.sqrt(1.0 - float(R(d.v0 * d.v0))))
    d.len_x_global = R(6.28318530717959 * _gam_sqrt_gam(gam))
    d.len_y_global = R(1.0)
    d.len_z_global = R(1.0)
    d.Npe = R(R(R(R(d.n0) * d.len_x_global) * d.len_y_global) * d.len_z_global)
    d.dt = R(0.99 * float(Deck.courant_length(R, d.len_x_global, d.len_y_global, d.len_z_global, d.nx, d.ny, d.nz))
             / float(R(d.c)))
    d.n0 = R(2.0)
    return d


def two_stream_short(real=np.float32, orientation="x") -> Deck:
    """decks/2stream-short.cxx:27-56 physics (v0=0.2, L=pi/2, nppc 100, 3000 steps).

    The deck as written pairs nx=32 with the y-oriented default initialiser and overruns the
    grid at HEAD.  orientation="x" keeps nx=32 and uses the x-oriented
    initialiser of decks/custom_init.cxx; orientation="y" is the historically run 1x32x1 form.
    """
    R = real
    if orientation == "x":
        d = Deck(real=real, name="2stream-short[x]", nx=32, ny=1, nz=1, init="custom_init")
        lens = (R(3.14159265358979 * 0.5), R(1.0), R(1.0))
    else:
        d = Deck(real=real, name="2stream-short[y]", nx=1, ny=32, nz=1, init="default")
        lens = (R(1.0), R(3.14159265358979 * 0.5), R(1.0))
    d.num_steps, d.nppc = 3000, 100
    d.v0 = R(0.2)
    d.len_x_global, d.len_y_global, d.len_z_global = lens
    d.Npe = R(R(R(R(d.n0) * d.len_x_global) * d.len_y_global) * d.len_z_global)
    d.dt = R(0.99 * float(Deck.courant_length(R, *lens, d.nx, d.ny, d.nz)) / float(R(d.c)))
    d.n0 = R(2.0)
    return d


def uniform_plasma(nx, ny, nz, nppc, real=np.float32, cell_size=0.1, num_steps=100) -> Deck:
    """Synthetic uniform thermal plasma: periodic box, dx = 0.1 d_e per axis,
    dt = 0.99*courant/c, E = cB = 0 initially, n0 = 1."""
    R = real
    d = Deck(real=real, name=f"uniform-{nx}x{ny}x{nz}x{nppc}", nx=nx, ny=ny, nz=nz, nppc=nppc,
             num_steps=num_steps, init="uniform")
    d.len_x_global, d.len_y_global, d.len_z_global = R(nx * cell_size), R(ny * cell_size), R(nz * cell_size)
    d.dt = R(0.99 * float(Deck.courant_length(R, d.len_x_global, d.len_y_global, d.len_z_global, nx, ny, nz))
             / float(R(d.c)))
    d.v0 = R(0.0)
    return d


def _gam_sqrt_gam(gam):
    """``gam*sqrt(gam)`` with ``real_ gam``: unqualified sqrt() on a float picks the C
    ``double sqrt(double)`` (only <cmath> is in scope), so the product is formed in double."""
    return float(gam) * math.sqrt(float(gam))

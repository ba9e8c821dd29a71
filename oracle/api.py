"""ctypes access to the parity oracle -- TEST INFRASTRUCTURE, not product code.

Only tests/, ``bench.py``'s cpu_baseline / ``--impl reference`` legs and
``__graft_entry__.smoke()`` may import this module; ``cabanapic_b200`` never does.

Two things live behind it:

* ``Restatement(prec)``  -- oracle/cpic_oracle.c, our plain-C restatement of the
  reference algorithm (built into oracle/build/ by oracle/Makefile);
* ``RefLib(deck, prec, omp)`` -- the reference's *own* sources compiled from
  /root/reference against include/compat/ (oracle/_ref/, built in the container
  and shipped prebuilt to the GPU box; may be absent -> ``RefLib.available``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "build")
REF = os.path.join(HERE, "_ref")

CONST_NAMES = "qdt_2mc cdt_dx cdt_dy cdt_dz qsp dx dy dz dt px py pz dt_eps0".split()
PARTICLE_NAMES = "dx dy dz ux uy uz w cell".split()
FIELD_NAMES = "ex ey ez cbx cby cbz jfx jfy jfz".split()
DECK_PARAM_NAMES = ("nx ny nz ng nppc num_steps num_particles num_cells dt c eps qsp me n0 Npe Ne v0 "
                    "len_x len_y len_z dx dy dz perform_uncenter").split()


class Consts(C.Structure):
    """Step constants as the driver derives them (example/example.cpp:77-113,179-181)."""
    _fields_ = [(n, C.c_double) for n in CONST_NAMES]

    @classmethod
    def from_dict(cls, d):
        return cls(**{n: float(d[n]) for n in CONST_NAMES})

    def to_dict(self):
        return {n: getattr(self, n) for n in CONST_NAMES}


def build(ref: bool = True) -> None:
    """(Re)build the oracle libraries with oracle/Makefile (idempotent)."""
    targets = ["restatement"] + (["ref"] if ref and os.path.isdir("/root/reference/src") else [])
    subprocess.run(["make", "-C", HERE, "-j8"] + targets, check=True, stdout=subprocess.DEVNULL)


def _dtype(prec: str):
    return {"f32": np.float32, "f64": np.float64}[prec]


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


class State:
    """Plain-array simulation state shared by every backend in the tests."""

    def __init__(self, nx, ny, nz, ng, np_, prec):
        self.nx, self.ny, self.nz, self.ng, self.np, self.prec = int(nx), int(ny), int(nz), int(ng), int(np_), prec
        self.nc = (nx + 2 * ng) * (ny + 2 * ng) * (nz + 2 * ng)
        r = _dtype(prec)
        self.p = {n: np.zeros(np_, dtype=(np.int32 if n == "cell" else r)) for n in PARTICLE_NAMES}
        self.f = np.zeros((9, self.nc), dtype=r)
        self.interp = np.zeros((self.nc, 18), dtype=r)
        self.acc = np.zeros((self.nc, 12), dtype=r)

    def copy(self):
        s = State(self.nx, self.ny, self.nz, self.ng, self.np, self.prec)
        for n in PARTICLE_NAMES:
            s.p[n][:] = self.p[n]
        s.f[:] = self.f
        s.interp[:] = self.interp
        s.acc[:] = self.acc
        return s


class Restatement:
    """oracle/cpic_oracle.c -- operates in place on a ``State``."""

    def __init__(self, prec: str = "f32"):
        path = os.path.join(BUILD, f"libcpic_oracle_{prec}.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        self.prec = prec
        assert self.lib.orc_real_bytes() == np.dtype(_dtype(prec)).itemsize
        self.lib.orc_push.restype = C.c_long

    @staticmethod
    def _grid(s):
        return C.c_long(s.nx), C.c_long(s.ny), C.c_long(s.nz), C.c_long(s.ng)

    def load_interpolator(self, s):
        self.lib.orc_load_interpolator(_ptr_array(list(s.f)), _ptr(s.interp), *self._grid(s))

    def clear_accumulator(self, s):
        self.lib.orc_clear_accumulator(_ptr(s.acc), C.c_long(s.nc))

    def push(self, s, k: Consts, periodic=True):
        ncross = C.c_long(0)
        movers = self.lib.orc_push(*[_ptr(s.p[n]) for n in PARTICLE_NAMES], C.c_long(s.np), _ptr(s.interp),
                                   _ptr(s.acc), C.byref(k), *self._grid(s),
                                   C.c_int(7 if periodic is True else int(periodic)),
                                   C.byref(ncross))
        return int(movers), int(ncross.value)

    def uncenter(self, s, qdt_2mc):
        p = s.p
        self.lib.orc_uncenter(_ptr(p["dx"]), _ptr(p["dy"]), _ptr(p["dz"]), _ptr(p["ux"]), _ptr(p["uy"]),
                              _ptr(p["uz"]), _ptr(p["cell"]), C.c_long(s.np), _ptr(s.interp), C.c_double(qdt_2mc))

    def unload_accumulator(self, s, k: Consts):
        self.lib.orc_unload_accumulator(_ptr_array(list(s.f)), _ptr(s.acc), *self._grid(s), C.byref(k))

    def advance_b(self, s, px, py, pz):
        self.lib.orc_advance_b(_ptr_array(list(s.f)), C.c_double(px), C.c_double(py), C.c_double(pz), *self._grid(s))

    def advance_e(self, s, px, py, pz, dt_eps0, solver=0):
        if solver == 0:
            self.lib.orc_advance_e_em(_ptr_array(list(s.f)), C.c_double(px), C.c_double(py), C.c_double(pz),
                                      *self._grid(s), C.c_double(dt_eps0))
        else:
            self.lib.orc_advance_e_es1d(_ptr_array(list(s.f)), *self._grid(s), C.c_double(dt_eps0))

    def ghost_copy(self, s, comps):
        self.lib.orc_ghost_copy(*[_ptr(s.f[c]) for c in comps], *self._grid(s))

    def ghost_fold(self, s, comps):
        self.lib.orc_ghost_fold(*[_ptr(s.f[c]) for c in comps], *self._grid(s))

    # slab-mode helpers (multi-GPU tests; see the comment in cpic_oracle.c)
    def ghost_copy_axes(self, s, comps, per):
        self.lib.orc_ghost_copy_axes(*[_ptr(s.f[c]) for c in comps], *self._grid(s), C.c_int(per))

    def ghost_fold_phase(self, s, phase, per):
        self.lib.orc_ghost_fold_phase(*[_ptr(s.f[c]) for c in (6, 7, 8)], *self._grid(s), C.c_int(phase), C.c_int(per))

    def advance_b_stencil(self, s, px, py, pz):
        self.lib.orc_advance_b_stencil(_ptr_array(list(s.f)), C.c_double(px), C.c_double(py), C.c_double(pz),
                                       *self._grid(s))

    def advance_e_stencil(self, s, px, py, pz, dt_eps0):
        self.lib.orc_advance_e_stencil(_ptr_array(list(s.f)), C.c_double(px), C.c_double(py), C.c_double(pz),
                                       *self._grid(s), C.c_double(dt_eps0))

    def pec_walls(self, s):
        """zero the tangential E on the six walls (Boundary::Reflect field side; unpinned by the reference)"""
        self.lib.orc_pec_walls(_ptr_array(list(s.f)), *self._grid(s))

    def step_reflect(self, s, k: Consts, nsteps=1):
        """the reference's loop (example/example.cpp:221-266) with reflecting particle walls and PEC field walls"""
        R = np.float32 if self.prec == "f32" else np.float64
        h = (float(R(0.5) * R(k.px)), float(R(0.5) * R(k.py)), float(R(0.5) * R(k.pz)))
        out = []
        for _ in range(nsteps):
            self.load_interpolator(s)
            self.clear_accumulator(s)
            self.push(s, k, periodic=7 << 4)
            self.unload_accumulator(s, k)
            self.advance_b_stencil(s, *h)
            self.advance_e_stencil(s, k.px, k.py, k.pz, k.dt_eps0)
            self.pec_walls(s)
            self.advance_b_stencil(s, *h)
            out.append(self.energies(s))
        return np.array(out)

    def energies(self, s, solver=0):
        e, b = C.c_double(), C.c_double()
        self.lib.orc_energies(_ptr_array(list(s.f)), C.c_int(solver), *self._grid(s), C.byref(e), C.byref(b))
        return e.value, b.value

    def step(self, s, k: Consts, solver=0, nsteps=1, energies=False):
        en = np.zeros(2 * nsteps) if energies else None
        self.lib.orc_step(*[_ptr(s.p[n]) for n in PARTICLE_NAMES], C.c_long(s.np), _ptr_array(list(s.f)),
                          _ptr(s.interp), _ptr(s.acc), C.byref(k), C.c_int(solver), *self._grid(s),
                          C.c_long(nsteps), _ptr(en) if energies else None)
        return en.reshape(nsteps, 2) if energies else None


class RefLib:
    """The reference's own hot-path sources (oracle/ref_driver.cpp + /root/reference/src)."""

    @staticmethod
    def path(deck="default", prec="f32", omp=False):
        return os.path.join(REF, f"libcpic_ref_{deck}_{prec}{'_omp' if omp else ''}.so")

    @classmethod
    def available(cls, deck="default", prec="f32", omp=False):
        return os.path.exists(cls.path(deck, prec, omp))

    def __init__(self, deck="default", prec="f32", omp=False):
        self.lib = L = C.CDLL(self.path(deck, prec, omp))
        self.prec, self.deck = prec, deck
        assert L.ref_real_bytes() == np.dtype(_dtype(prec)).itemsize
        L.ref_create.restype = C.c_void_p
        L.ref_create_from_deck.restype = C.c_void_p
        L.ref_num_cells.restype = C.c_long
        L.ref_num_particles.restype = C.c_long
        self.h = None
        self.solver = 0

    def num_threads(self):
        return int(self.lib.ref_num_threads())

    # -- deck seam -----------------------------------------------------
    def deck_params(self):
        out = (C.c_double * 24)()
        self.lib.ref_deck_params(out)
        d = dict(zip(DECK_PARAM_NAMES, list(out)))
        for n in "nx ny nz ng nppc num_steps num_particles num_cells".split():
            d[n] = int(d[n])
        return d

    def deck_consts(self):
        k, dxp, we = Consts(), C.c_double(), C.c_double()
        self.lib.ref_deck_consts(C.byref(k), C.byref(dxp), C.byref(we))
        return k, dxp.value, we.value

    def create_from_deck(self, solver=0):
        self.destroy()
        self.solver = solver
        self.h = C.c_void_p(self.lib.ref_create_from_deck(C.c_int(solver)))
        return self

    # -- explicit state ---------------------------------------------------
    def create(self, s: State, solver=0):
        self.destroy()
        self.solver = solver
        self.h = C.c_void_p(self.lib.ref_create(C.c_long(s.nx), C.c_long(s.ny), C.c_long(s.nz), C.c_long(s.ng),
                                                C.c_long(s.np), C.c_int(solver)))
        self.put(s)
        return self

    def destroy(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def put(self, s: State):
        self.lib.ref_set_particles(self.h, *[_ptr(s.p[n]) for n in PARTICLE_NAMES])
        self.lib.ref_set_fields(self.h, _ptr_array(list(s.f)))
        self.lib.ref_set_interpolators(self.h, _ptr(s.interp))
        self.lib.ref_set_accumulators(self.h, _ptr(s.acc))

    def get(self, s: State | None = None, grid=None):
        if s is None:
            nx, ny, nz, ng = grid
            s = State(nx, ny, nz, ng, int(self.lib.ref_num_particles(self.h)), self.prec)
        self.lib.ref_get_particles(self.h, *[_ptr(s.p[n]) for n in PARTICLE_NAMES])
        self.lib.ref_get_fields(self.h, _ptr_array(list(s.f)))
        self.lib.ref_get_interpolators(self.h, _ptr(s.interp))
        self.lib.ref_get_accumulators(self.h, _ptr(s.acc))
        return s

    # -- the reference's entry points --------------------------------------
    def load_interpolator(self):
        self.lib.ref_load_interpolator(self.h)

    def clear_accumulator(self):
        self.lib.ref_clear_accumulator(self.h)

    def push(self, k: Consts):
        self.lib.ref_push(self.h, C.byref(k))

    def unload_accumulator(self, k: Consts):
        self.lib.ref_unload_accumulator(self.h, C.byref(k))

    def advance_b(self, px, py, pz):
        self.lib.ref_advance_b(self.h, C.c_double(px), C.c_double(py), C.c_double(pz))

    def advance_e(self, px, py, pz, dt_eps0):
        self.lib.ref_advance_e(self.h, C.c_double(px), C.c_double(py), C.c_double(pz), C.c_double(dt_eps0))

    def uncenter(self, qdt_2mc):
        self.lib.ref_uncenter(self.h, C.c_double(qdt_2mc))

    def energies(self):
        e, b = C.c_double(), C.c_double()
        self.lib.ref_energies(self.h, C.byref(e), C.byref(b))
        return e.value, b.value

    def run(self, k: Consts, nsteps=1, energies=False):
        en = np.zeros(2 * nsteps) if energies else None
        self.lib.ref_run(self.h, C.byref(k), C.c_long(nsteps), _ptr(en) if energies else None)
        return en.reshape(nsteps, 2) if energies else None

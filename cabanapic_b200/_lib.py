"""ctypes binding of the C ABI in include/cabanapic_b200.h.

This is plumbing for tests and bench.py: the product is the shared library
(``cabanapic_b200/libcabanapic_b200.so``, built from ``csrc/`` for sm_100a) and
its C++ host facade (``include/cabanapic/``).  Nothing here computes anything;
if the library is missing or no CUDA device is usable we fail loudly -- there
is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CPIC_LIB") or os.path.join(HERE, "libcabanapic_b200.so")   # CPIC_LIB: developer override

CONST_NAMES = "qdt_2mc cdt_dx cdt_dy cdt_dz qsp dx dy dz dt px py pz dt_eps0".split()
PARTICLE_NAMES = "dx dy dz ux uy uz w cell".split()
FIELD_NAMES = "ex ey ez cbx cby cbz jfx jfy jfz".split()

SOLVER_EM, SOLVER_ES_1D = 0, 1
BOUNDARY_REFLECT, BOUNDARY_PERIODIC = 0, 1
FP_STRICT, FP_CONTRACT = 0, 1
DEPOSIT_AUTO, DEPOSIT_ATOMIC, DEPOSIT_ATOMIC_V4, DEPOSIT_WARP, DEPOSIT_ORDERED = 0, 1, 2, 3, 4
SORT_FUSED = -1      # cpic_step sort_interval: keep the store cell-ordered with the reordering push

ERROR_NAMES = {-1: "CPIC_E_INVALID", -2: "CPIC_E_CUDA", -3: "CPIC_E_NOMEM", -4: "CPIC_E_CAPACITY",
               -5: "CPIC_E_BAD_CELL", -6: "CPIC_E_UNSUPPORTED"}


class CpicError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("ng", C.c_int32),
                ("real_bytes", C.c_int32), ("solver", C.c_int32), ("boundary", C.c_int32), ("device", C.c_int32),
                ("fp_mode", C.c_int32), ("deposit_mode", C.c_int32), ("max_particles", C.c_int64),
                ("enable_sort", C.c_int32), ("reserved", C.c_int32 * 7)]


class Consts(C.Structure):
    _fields_ = [(n, C.c_double) for n in CONST_NAMES]

    @classmethod
    def from_dict(cls, d):
        return cls(**{n: float(d[n]) for n in CONST_NAMES})

    def to_dict(self):
        return {n: getattr(self, n) for n in CONST_NAMES}


class PushStats(C.Structure):
    _fields_ = [("movers", C.c_int64), ("crossings", C.c_int64), ("wraps", C.c_int64 * 6)]


def build(verbose: bool = False) -> str:
    """Compile csrc/ into the in-tree shared library (nvcc, sm_100a)."""
    r = subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libcabanapic_b200.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """The loaded C-ABI library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no fallback implementation)")
        L = C.CDLL(LIB_PATH)
        L.cpic_last_error.restype = C.c_char_p
        L.cpic_last_error.argtypes = [C.c_void_p]
        L.cpic_destroy.restype = None
        L.cpic_destroy.argtypes = [C.c_void_p]
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        sig = {
            "cpic_abi_version": [],
            "cpic_create": [C.POINTER(Params), C.POINTER(vp)],
            "cpic_sync": [vp],
            "cpic_num_cells": [vp, C.POINTER(i64)],
            "cpic_num_particles": [vp, C.POINTER(i64)],
            "cpic_upload_particles": [vp] + [vp] * 8 + [i64],
            "cpic_download_particles": [vp] + [vp] * 8 + [i64, C.POINTER(i64)],
            "cpic_upload_fields": [vp, C.POINTER(vp)],
            "cpic_download_fields": [vp, C.POINTER(vp)],
            "cpic_upload_interpolators": [vp, vp],
            "cpic_download_interpolators": [vp, vp],
            "cpic_upload_accumulators": [vp, vp],
            "cpic_download_accumulators": [vp, vp],
            "cpic_load_interpolator_array": [vp],
            "cpic_initialize_interpolator": [vp],
            "cpic_clear_accumulator_array": [vp],
            "cpic_push": [vp, C.POINTER(Consts)],
            "cpic_contribute": [vp],
            "cpic_unload_accumulator_array": [vp, C.POINTER(Consts)],
            "cpic_advance_b": [vp, dbl, dbl, dbl],
            "cpic_advance_e": [vp, dbl, dbl, dbl, dbl],
            "cpic_uncenter_particles": [vp, dbl],
            "cpic_energies": [vp, C.POINTER(dbl), C.POINTER(dbl)],
            "cpic_kinetic_energy": [vp, C.POINTER(dbl)],
            "cpic_update_ghosts": [vp, C.c_int],
            "cpic_step": [vp, C.POINTER(Consts), i64, i32, vp],
            "cpic_step_host": [vp, C.POINTER(Consts), C.POINTER(vp), C.POINTER(vp), i64, C.POINTER(vp), C.POINTER(vp), vp],
            "cpic_sort_particles": [vp],
            "cpic_push_reorder": [vp, C.POINTER(Consts)],
            "cpic_init_uniform_plasma": [vp, i64, i64, i32, i32, i32, i32, i32, C.c_uint64, dbl, dbl, dbl, dbl],
            "cpic_enable_push_stats": [vp, i32],
            "cpic_push_stats_get": [vp, C.POINTER(PushStats)],
            "cpic_device_ptr": [vp, C.c_int, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64)],
            "cpic_set_stream": [vp, vp],
            "cpic_set_num_particles": [vp, i64],
            "cpic_set_modes": [vp, i32, i32],
            "cpic_set_axis_periodic": [vp, i32, i32, i32],
            "cpic_advance_b_stencil": [vp, dbl, dbl, dbl],
            "cpic_advance_e_stencil": [vp, dbl, dbl, dbl, dbl],
            "cpic_extract_z_leavers": [vp, vp, vp, i64, C.POINTER(i64), C.POINTER(i64), i32, i32],
            "cpic_append_particles_device": [vp, vp, i64, i64],
            "cpic_slab_extract_async": [vp, vp, vp, i64, vp, i32, i32],
            "cpic_slab_append_async": [vp, vp, i64, vp],
            "cpic_last_ms": [vp, C.c_int, C.POINTER(dbl)],
            "cpic_launch_count": [vp, C.POINTER(i64)],
            "cpic_enable_step_profile": [vp, i32],
            "cpic_step_profile": [vp, C.POINTER(dbl), C.POINTER(i64)],
            "cpic_state_digest": [vp, C.POINTER(dbl)],
            "cpic_create_species": [vp, i64, C.POINTER(vp)],
            "cpic_step_species": [vp, C.POINTER(vp), C.POINTER(Consts), i32, i64, i32, vp],
            # multi-GPU layer (include/cabanapic_b200_mgpu.h)
            "cpic_mgpu_unique_id": [vp],
            "cpic_mgpu_bootstrap_file": [C.c_char_p, i32, i32, dbl, vp],
            "cpic_mgpu_create": [C.POINTER(Params), i32, i32, vp, i32, i64, C.POINTER(vp)],
            "cpic_mgpu_layout": [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
            "cpic_mgpu_init_uniform_plasma": [vp, i32, C.c_uint64, dbl, dbl, dbl, dbl],
            "cpic_mgpu_reduce_accumulator": [vp],
            "cpic_mgpu_step": [vp, C.POINTER(Consts), i64, i32, i32],
            "cpic_mgpu_prepare_graph": [vp, C.POINTER(Consts)],
            "cpic_mgpu_migration_counts": [vp, C.POINTER(i64)],
            "cpic_mgpu_last_migration": [vp, C.POINTER(i64)],
            "cpic_mgpu_energies": [vp, C.POINTER(dbl), C.POINTER(dbl)],
            "cpic_mgpu_state_digest": [vp, C.POINTER(dbl)],
            "cpic_mgpu_sync": [vp],
            "cpic_mgpu_used_graph": [vp],
            "cpic_mgpu_transport": [vp],
            "cpic_mgpu_step_host": [vp, C.POINTER(Consts), C.POINTER(vp), C.POINTER(vp), i64, i64, C.POINTER(i64), C.POINTER(vp), C.POINTER(vp)],
        }
        L.cpic_mgpu_last_error.restype = C.c_char_p
        L.cpic_mgpu_last_error.argtypes = [C.c_void_p]
        L.cpic_mgpu_destroy.restype = None
        L.cpic_mgpu_destroy.argtypes = [C.c_void_p]
        L.cpic_mgpu_context.restype = C.c_void_p
        L.cpic_mgpu_context.argtypes = [C.c_void_p]
        for name, args in sig.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = L
    return _lib


EXPORTED = ["cpic_abi_version", "cpic_last_error", "cpic_create", "cpic_destroy", "cpic_sync", "cpic_num_cells",
            "cpic_num_particles", "cpic_upload_particles", "cpic_download_particles", "cpic_upload_fields",
            "cpic_download_fields", "cpic_upload_interpolators", "cpic_download_interpolators",
            "cpic_upload_accumulators", "cpic_download_accumulators", "cpic_load_interpolator_array",
            "cpic_initialize_interpolator", "cpic_clear_accumulator_array", "cpic_push", "cpic_contribute",
            "cpic_unload_accumulator_array", "cpic_advance_b", "cpic_advance_e", "cpic_uncenter_particles",
            "cpic_energies", "cpic_kinetic_energy", "cpic_update_ghosts", "cpic_step", "cpic_step_host", "cpic_sort_particles", "cpic_push_reorder", "cpic_init_uniform_plasma", "cpic_enable_push_stats",
            "cpic_push_stats_get", "cpic_device_ptr", "cpic_set_stream", "cpic_set_num_particles", "cpic_set_modes",
            "cpic_set_axis_periodic", "cpic_advance_b_stencil", "cpic_advance_e_stencil", "cpic_extract_z_leavers", "cpic_append_particles_device", "cpic_slab_extract_async", "cpic_slab_append_async",
            "cpic_last_ms", "cpic_launch_count", "cpic_enable_step_profile", "cpic_step_profile", "cpic_state_digest",
            "cpic_create_species", "cpic_step_species"]
EXPORTED_MGPU = ["cpic_mgpu_last_error", "cpic_mgpu_unique_id", "cpic_mgpu_bootstrap_file", "cpic_mgpu_create",
                 "cpic_mgpu_destroy", "cpic_mgpu_context", "cpic_mgpu_layout", "cpic_mgpu_init_uniform_plasma",
                 "cpic_mgpu_reduce_accumulator", "cpic_mgpu_step", "cpic_mgpu_prepare_graph", "cpic_mgpu_migration_counts", "cpic_mgpu_last_migration",
                 "cpic_mgpu_energies", "cpic_mgpu_state_digest", "cpic_mgpu_sync", "cpic_mgpu_used_graph", "cpic_mgpu_transport", "cpic_mgpu_step_host"]
TRANSPORT_NAMES = {0: "none", 1: "nccl", 2: "peer-memory"}
MGPU_REPLICATED, MGPU_SLAB, MGPU_AUTO = 0, 1, 2
DIGEST_NAMES = ["particles", "weight_sum", "cells_not_interior", "offsets_out_of_range", "kinetic_energy",
                "e_energy", "b_energy", "migrated"]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One simulation context on one GPU (thin wrapper over ``cpic_ctx``).

    Method names follow the reference's call surface (example/example.cpp:221-266).
    All arrays are host numpy arrays; ``real`` is float32 or float64.
    """

    def __init__(self, nx, ny, nz, ng=1, max_particles=0, real=np.float32, solver=SOLVER_EM,
                 boundary=BOUNDARY_PERIODIC, device=0, fp_mode=FP_STRICT, deposit_mode=DEPOSIT_AUTO,
                 enable_sort=True):
        self.L = lib()
        self.real = np.dtype(real)
        self.params = Params(nx=nx, ny=ny, nz=nz, ng=ng, real_bytes=self.real.itemsize, solver=solver,
                             boundary=boundary, device=device, fp_mode=fp_mode, deposit_mode=deposit_mode,
                             max_particles=int(max_particles), enable_sort=1 if enable_sort else 0)
        h = C.c_void_p()
        rc = self.L.cpic_create(C.byref(self.params), C.byref(h))
        if rc != 0:
            raise CpicError(rc, (self.L.cpic_last_error(None) or b"").decode())
        self.h = h
        self.nx, self.ny, self.nz, self.ng = nx, ny, nz, ng
        self.nc = (nx + 2 * ng) * (ny + 2 * ng) * (nz + 2 * ng)
        self.solver = solver

    @classmethod
    def borrowed(cls, handle, nx, ny, nz, real, solver=SOLVER_EM):
        """Wrap a cpic_ctx owned by somebody else (cpic_mgpu_context): never destroyed from here."""
        self = cls.__new__(cls)
        self.L = lib()
        self.real = np.dtype(real)
        self.h = C.c_void_p(handle)
        self._borrowed = True
        self.nx, self.ny, self.nz, self.ng = nx, ny, nz, 1
        self.nc = (nx + 2) * (ny + 2) * (nz + 2)
        self.solver = solver
        return self

    def close(self):
        if getattr(self, "h", None):
            if not getattr(self, "_borrowed", False):
                self.L.cpic_destroy(self.h)
            self.h = None

    def state_digest(self):
        d = (C.c_double * 8)()
        self._ck(self.L.cpic_state_digest(self.h, d))
        return dict(zip(DIGEST_NAMES, list(d)))

    def create_species(self, max_particles):
        """another particle store (species) on this context's fields / accumulators; close it before this context"""
        h = C.c_void_p()
        self._ck(self.L.cpic_create_species(self.h, int(max_particles), C.byref(h)))
        sp = Context.borrowed(h.value, self.nx, self.ny, self.nz, self.real, self.solver)
        sp._borrowed = False          # owned: cpic_destroy frees only what the species itself allocated
        sp._parent = self
        return sp

    def step_species(self, species, consts, nsteps=1, sort_interval=0, energies=False):
        n = len(species)
        hs = (C.c_void_p * n)(*[s.h.value if isinstance(s.h, C.c_void_p) else s.h for s in species])
        ks = (Consts * n)(*consts)
        en = np.zeros((nsteps, 2), dtype=np.float64) if energies else None
        self._ck(self.L.cpic_step_species(self.h, hs, ks, n, nsteps, sort_interval, _p(en)))
        return en

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise CpicError(rc, (self.L.cpic_last_error(self.h) or b"").decode())

    # ---- transfers ------------------------------------------------------------------
    @property
    def num_particles(self):
        n = C.c_int64()
        self._ck(self.L.cpic_num_particles(self.h, C.byref(n)))
        return n.value

    def upload_particles(self, p: dict):
        n = len(p["cell"])
        arrs = [np.ascontiguousarray(p[k], dtype=self.real) for k in PARTICLE_NAMES[:7]]
        cell = np.ascontiguousarray(p["cell"], dtype=np.int32)
        self._ck(self.L.cpic_upload_particles(self.h, *[_p(a) for a in arrs], _p(cell), n))

    def download_particles(self) -> dict:
        n = self.num_particles
        out = {k: np.empty(n, dtype=self.real) for k in PARTICLE_NAMES[:7]}
        out["cell"] = np.empty(n, dtype=np.int32)
        got = C.c_int64()
        self._ck(self.L.cpic_download_particles(self.h, *[_p(out[k]) for k in PARTICLE_NAMES], n, C.byref(got)))
        return out

    def upload_fields(self, f):
        f = np.ascontiguousarray(f, dtype=self.real)
        assert f.shape == (9, self.nc)
        ptrs = (C.c_void_p * 9)(*[f[m].ctypes.data for m in range(9)])
        self._ck(self.L.cpic_upload_fields(self.h, ptrs))

    def download_fields(self):
        f = np.empty((9, self.nc), dtype=self.real)
        ptrs = (C.c_void_p * 9)(*[f[m].ctypes.data for m in range(9)])
        self._ck(self.L.cpic_download_fields(self.h, ptrs))
        return f

    def upload_interpolators(self, a):
        a = np.ascontiguousarray(a, dtype=self.real)
        assert a.shape == (self.nc, 18)
        self._ck(self.L.cpic_upload_interpolators(self.h, _p(a)))

    def download_interpolators(self):
        a = np.empty((self.nc, 18), dtype=self.real)
        self._ck(self.L.cpic_download_interpolators(self.h, _p(a)))
        return a

    def upload_accumulators(self, a):
        a = np.ascontiguousarray(a, dtype=self.real)
        assert a.shape == (self.nc, 12)
        self._ck(self.L.cpic_upload_accumulators(self.h, _p(a)))

    def download_accumulators(self):
        a = np.empty((self.nc, 12), dtype=self.real)
        self._ck(self.L.cpic_download_accumulators(self.h, _p(a)))
        return a

    # ---- the time-loop call surface ---------------------------------------------------
    def load_interpolator_array(self):
        self._ck(self.L.cpic_load_interpolator_array(self.h))

    def initialize_interpolator(self):
        self._ck(self.L.cpic_initialize_interpolator(self.h))

    def clear_accumulator_array(self):
        self._ck(self.L.cpic_clear_accumulator_array(self.h))

    def push(self, k: Consts):
        self._ck(self.L.cpic_push(self.h, C.byref(k)))

    def contribute(self):
        self._ck(self.L.cpic_contribute(self.h))

    def unload_accumulator_array(self, k: Consts):
        self._ck(self.L.cpic_unload_accumulator_array(self.h, C.byref(k)))

    def advance_b(self, px, py, pz):
        self._ck(self.L.cpic_advance_b(self.h, px, py, pz))

    def advance_e(self, px, py, pz, dt_eps0):
        self._ck(self.L.cpic_advance_e(self.h, px, py, pz, dt_eps0))

    def uncenter_particles(self, qdt_2mc):
        self._ck(self.L.cpic_uncenter_particles(self.h, qdt_2mc))

    def update_ghosts(self, which):
        self._ck(self.L.cpic_update_ghosts(self.h, which))

    def energies(self):
        e, b = C.c_double(), C.c_double()
        self._ck(self.L.cpic_energies(self.h, C.byref(e), C.byref(b)))
        return e.value, b.value

    def kinetic_energy(self):
        """sum_p w_p (gamma_p - 1), in double (cpic_kinetic_energy)"""
        v = C.c_double()
        self._ck(self.L.cpic_kinetic_energy(self.h, C.byref(v)))
        return v.value

    def step(self, k: Consts, nsteps=1, sort_interval=0, energies=False):
        en = np.zeros((nsteps, 2)) if energies else None
        self._ck(self.L.cpic_step(self.h, C.byref(k), nsteps, sort_interval, _p(en)))
        return en

    def step_host(self, k: Consts, p_in: dict, p_out: dict | None, f_in, f_out=None, energies=False):
        """cpic_step_host: one step on HOST-resident state, particles streamed through the device in chunks
        (H2D, push, D2H overlapped).  p_in / p_out: dicts of the eight member arrays (p_out may be p_in);
        f_in / f_out: (9, nc) arrays.  Arrays are used as they are (no copies): they must be contiguous and of
        the context's real type / int32; pinned memory makes the transfers asynchronous."""
        n = len(p_in["cell"])
        for d in (p_in, p_out):
            if d is not None:
                for kk in PARTICLE_NAMES:
                    a = d[kk]
                    want = np.int32 if kk == "cell" else self.real
                    assert a.dtype == want and a.flags.c_contiguous and len(a) >= n, kk
        pin = (C.c_void_p * 8)(*[p_in[kk].ctypes.data for kk in PARTICLE_NAMES])
        pout = None if p_out is None else (C.c_void_p * 8)(*[p_out[kk].ctypes.data for kk in PARTICLE_NAMES])
        assert f_in.dtype == self.real and f_in.shape == (9, self.nc) and f_in.flags.c_contiguous
        fin = (C.c_void_p * 9)(*[f_in[m].ctypes.data for m in range(9)])
        fout = None
        if f_out is not None:
            assert f_out.dtype == self.real and f_out.shape == (9, self.nc) and f_out.flags.c_contiguous
            fout = (C.c_void_p * 9)(*[f_out[m].ctypes.data for m in range(9)])
        en = np.zeros(2) if energies else None
        self._ck(self.L.cpic_step_host(self.h, C.byref(k), pin, pout, n, fin, fout, _p(en)))
        return en

    def sort_particles(self):
        self._ck(self.L.cpic_sort_particles(self.h))

    def push_reorder(self, k: Consts):
        """cpic_push + the cell ordering of the store in one pass (include/cabanapic_b200.h)."""
        self._ck(self.L.cpic_push_reorder(self.h, C.byref(k)))

    def init_uniform_plasma(self, first, count, gnx, gny, gnz, nppc, z0=0, seed=12345, vth=(0.1, 0.1, 0.1),
                            weight=1.0):
        self._ck(self.L.cpic_init_uniform_plasma(self.h, first, count, gnx, gny, gnz, nppc, z0, seed,
                                                 float(vth[0]), float(vth[1]), float(vth[2]), float(weight)))

    def sync(self):
        self._ck(self.L.cpic_sync(self.h))

    # ---- diagnostics / interop ----------------------------------------------------------
    def enable_push_stats(self, on=True):
        self._ck(self.L.cpic_enable_push_stats(self.h, 1 if on else 0))

    def push_stats(self):
        s = PushStats()
        self._ck(self.L.cpic_push_stats_get(self.h, C.byref(s)))
        return {"movers": s.movers, "crossings": s.crossings, "wraps": list(s.wraps)}

    def set_modes(self, fp_mode, deposit_mode):
        self._ck(self.L.cpic_set_modes(self.h, fp_mode, deposit_mode))

    def set_num_particles(self, n):
        self._ck(self.L.cpic_set_num_particles(self.h, n))

    def set_axis_periodic(self, px, py, pz):
        self._ck(self.L.cpic_set_axis_periodic(self.h, int(px), int(py), int(pz)))

    def advance_b_stencil(self, px, py, pz):
        self._ck(self.L.cpic_advance_b_stencil(self.h, px, py, pz))

    def advance_e_stencil(self, px, py, pz, dt_eps0):
        self._ck(self.L.cpic_advance_e_stencil(self.h, px, py, pz, dt_eps0))

    def extract_z_leavers(self, lo_ptr, hi_ptr, capacity, rebase_lo, rebase_hi):
        """Device pointers in, (n_lo, n_hi) out; see include/cabanapic_b200.h."""
        a, b = C.c_int64(), C.c_int64()
        self._ck(self.L.cpic_extract_z_leavers(self.h, C.c_void_p(lo_ptr), C.c_void_p(hi_ptr), capacity, C.byref(a),
                                               C.byref(b), rebase_lo, rebase_hi))
        return a.value, b.value

    def slab_extract_async(self, lo_ptr, hi_ptr, capacity, counts_dev_ptr, rebase_lo, rebase_hi):
        """cpic_slab_extract_async: counts (int64[2], device) instead of a host round trip."""
        self._ck(self.L.cpic_slab_extract_async(self.h, C.c_void_p(lo_ptr), C.c_void_p(hi_ptr), capacity,
                                                C.c_void_p(counts_dev_ptr), rebase_lo, rebase_hi))

    def slab_append_async(self, ptr, capacity, count_dev_ptr):
        self._ck(self.L.cpic_slab_append_async(self.h, C.c_void_p(ptr), capacity, C.c_void_p(count_dev_ptr)))

    def append_particles_device(self, ptr, capacity, n):
        self._ck(self.L.cpic_append_particles_device(self.h, C.c_void_p(ptr), capacity, n))

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.L.cpic_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def device_ptr(self, which):
        p, n, s = C.c_void_p(), C.c_int64(), C.c_int64()
        self._ck(self.L.cpic_device_ptr(self.h, which, C.byref(p), C.byref(n), C.byref(s)))
        return p.value, n.value, s.value

    def last_ms(self, what):
        ms = C.c_double()
        self._ck(self.L.cpic_last_ms(self.h, what, C.byref(ms)))
        return ms.value

    def enable_step_profile(self, on=True):
        self._ck(self.L.cpic_enable_step_profile(self.h, 1 if on else 0))

    def step_profile(self):
        """ms spent in (sort, interp+clear, push, field side) over the last profiled step() call."""
        ms = (C.c_double * 4)()
        n = C.c_int64()
        self._ck(self.L.cpic_step_profile(self.h, ms, C.byref(n)))
        return {"sort_ms": ms[0], "interp_ms": ms[1], "push_ms": ms[2], "field_ms": ms[3], "steps": n.value}

    @property
    def launch_count(self):
        n = C.c_int64()
        self._ck(self.L.cpic_launch_count(self.h, C.byref(n)))
        return n.value


class Mgpu:
    """One rank of the native multi-GPU layer (include/cabanapic_b200_mgpu.h): host C++ + NCCL inside the library.
    The launcher only has to distribute the 128-byte NCCL id (``unique_id()`` on rank 0)."""

    def __init__(self, nx, ny, nz, rank, world, unique_id: bytes | None, mode=MGPU_AUTO, max_particles=0, real=np.float32,
                 solver=SOLVER_EM, device=0, fp_mode=FP_STRICT, send_capacity=0):
        self.L = lib()
        self.real = np.dtype(real)
        self.params = Params(nx=nx, ny=ny, nz=nz, ng=1, real_bytes=self.real.itemsize, solver=solver,
                             boundary=BOUNDARY_PERIODIC, device=device, fp_mode=fp_mode, deposit_mode=DEPOSIT_AUTO,
                             max_particles=int(max_particles), enable_sort=1)
        h = C.c_void_p()
        idbuf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        rc = self.L.cpic_mgpu_create(C.byref(self.params), rank, world, idbuf, mode, int(send_capacity), C.byref(h))
        if rc != 0:
            raise CpicError(rc, (self.L.cpic_mgpu_last_error(None) or b"").decode())
        self.h = h
        self.rank, self.world = rank, world
        mo, z0, nzl = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self.L.cpic_mgpu_layout(self.h, C.byref(mo), C.byref(z0), C.byref(nzl)))
        self.mode, self.z0, self.nzl = mo.value, z0.value, nzl.value
        self.ctx = Context.borrowed(self.L.cpic_mgpu_context(self.h), nx, ny, self.nzl, real, solver)

    @staticmethod
    def unique_id() -> bytes:
        L = lib()
        buf = C.create_string_buffer(128)
        rc = L.cpic_mgpu_unique_id(buf)
        if rc != 0:
            raise CpicError(rc, (L.cpic_mgpu_last_error(None) or b"").decode())
        return buf.raw

    def _ck(self, rc):
        if rc != 0:
            raise CpicError(rc, (self.L.cpic_mgpu_last_error(self.h) or b"").decode())

    def init_uniform_plasma(self, nppc, seed=12345, vth=(0.1, 0.1, 0.1), weight=1.0):
        self._ck(self.L.cpic_mgpu_init_uniform_plasma(self.h, nppc, seed, vth[0], vth[1], vth[2], weight))

    def step(self, k: Consts, nsteps=1, sort_interval=SORT_FUSED, use_graph=False):
        self._ck(self.L.cpic_mgpu_step(self.h, C.byref(k), nsteps, sort_interval, 1 if use_graph else 0))

    def step_host(self, k: Consts, p_in, p_out, n, fields_in, fields_out):
        """cpic_mgpu_step_host: p_in / p_out are dicts of host member arrays (p_out's at least as long as the count after
        the step), fields_* (9, nc) host arrays; returns this rank's particle count after the step."""
        names = "dx dy dz ux uy uz w cell".split()
        pin = (C.c_void_p * 8)(*[p_in[m].ctypes.data for m in names])
        pout = (C.c_void_p * 8)(*[p_out[m].ctypes.data for m in names])
        fin = (C.c_void_p * 9)(*[fields_in[m].ctypes.data for m in range(9)])
        fout = (C.c_void_p * 9)(*[fields_out[m].ctypes.data for m in range(9)])
        cap = min(len(p_out[m]) for m in names)
        n_out = C.c_int64()
        self._ck(self.L.cpic_mgpu_step_host(self.h, C.byref(k), pin, pout, int(n), int(cap), C.byref(n_out), fin, fout))
        return int(n_out.value)

    def prepare_graph(self, k: Consts):
        """capture the pair-of-steps graph now (nothing executes); False when the graph path does not apply"""
        return self.L.cpic_mgpu_prepare_graph(self.h, C.byref(k)) == 0

    def reduce_accumulator(self):
        self._ck(self.L.cpic_mgpu_reduce_accumulator(self.h))

    def migration_counts(self):
        a = (C.c_int64 * 2)()
        self._ck(self.L.cpic_mgpu_migration_counts(self.h, a))
        return int(a[0]), int(a[1])

    def last_migration(self):
        a = (C.c_int64 * 2)()
        self._ck(self.L.cpic_mgpu_last_migration(self.h, a))
        return int(a[0]), int(a[1])

    def energies(self):
        e, b = C.c_double(), C.c_double()
        self._ck(self.L.cpic_mgpu_energies(self.h, C.byref(e), C.byref(b)))
        return e.value, b.value

    def state_digest(self):
        d = (C.c_double * 8)()
        self._ck(self.L.cpic_mgpu_state_digest(self.h, d))
        return dict(zip(DIGEST_NAMES, list(d)))

    def sync(self):
        self._ck(self.L.cpic_mgpu_sync(self.h))

    @property
    def used_graph(self):
        return bool(self.L.cpic_mgpu_used_graph(self.h))

    @property
    def transport(self):
        """how the slab exchanges travel: 'peer-memory' (NVLink stores into the neighbours' memory), 'nccl' or 'none'"""
        return TRANSPORT_NAMES[self.L.cpic_mgpu_transport(self.h)]

    def close(self):
        if getattr(self, "h", None):
            self.ctx.close()
            self.L.cpic_mgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

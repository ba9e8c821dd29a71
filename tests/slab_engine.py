"""TEST INFRASTRUCTURE: an oracle-backed engine with the interface cabanapic_b200.dist expects, so
the slab / replicated exchange choreography can run on CPU under gloo (world_size 2).  The compute
is oracle/cpic_oracle.c; the product's GpuEngine is exercised by the `-m gpu` tests instead."""
import numpy as np
import torch

from oracle.api import PARTICLE_NAMES, Consts as OConsts, Restatement, State


class OracleEngine:
    def __init__(self, nx, ny, nz, capacity, prec="f32", z_periodic=True, async_migration=False):
        # async_migration: offer the device-counted migration interface of GpuEngine (counts travel as tensors, the
        # stepper never sees them as Python ints) so that its exchange choreography runs under gloo too
        self.async_migration = async_migration
        self.nx, self.ny, self.nz = nx, ny, nz
        self.prec = prec
        self.real = np.dtype(np.float32 if prec == "f32" else np.float64)
        self.gx, self.gy, self.gz = nx + 2, ny + 2, nz + 2
        self.plane = self.gx * self.gy
        self.nc = self.plane * self.gz
        self.O = Restatement(prec)
        self.cap = capacity
        self.s = State(nx, ny, nz, 1, capacity, prec)     # arrays sized to the capacity; s.np = live count
        self.s.np = 0
        self.per = 7 if z_periodic else 3
        self._f = torch.from_numpy(self.s.f)
        self._acc = torch.from_numpy(self.s.acc).view(self.gz, self.plane * 12)

    def set_particles(self, p):
        n = len(p["cell"])
        for k in PARTICLE_NAMES:
            self.s.p[k][:n] = p[k]
        self.s.np = n

    def particles(self):
        return {k: self.s.p[k][:self.s.np].copy() for k in PARTICLE_NAMES}

    def _ok(self, k):
        return OConsts(**k.to_dict()) if not isinstance(k, OConsts) else k

    def field_planes(self, m): return self._f[m].view(self.gz, self.plane)
    def acc_planes(self): return self._acc
    def load_interpolator(self): self.O.load_interpolator(self.s)
    def clear_accumulator(self): self.O.clear_accumulator(self.s)
    def push(self, k): self.O.push(self.s, self._ok(k), periodic=self.per)
    def push_reorder(self, k): self.push(k)      # the order of the store is not part of the oracle's state
    def unload_accumulator(self, k): self.O.unload_accumulator(self.s, self._ok(k))
    def fold_phase(self, phase): self.O.ghost_fold_phase(self.s, phase, self.per)
    def ghost_copy_local(self, which): self.O.ghost_copy_axes(self.s, (6, 7, 8) if which == "J" else (3, 4, 5), self.per)
    def advance_b_stencil(self, px, py, pz): self.O.advance_b_stencil(self.s, px, py, pz)
    def advance_e_stencil(self, px, py, pz, cj): self.O.advance_e_stencil(self.s, px, py, pz, cj)
    def advance_b(self, px, py, pz): self.O.advance_b(self.s, px, py, pz)
    def advance_e(self, px, py, pz, cj): self.O.advance_e(self.s, px, py, pz, cj)
    def sort(self): pass
    def energies(self): return self.O.energies(self.s, 0)
    def sync(self): pass
    def close(self): pass

    @property
    def num_particles(self): return self.s.np

    def alloc_bytes(self, n): return torch.empty(max(int(n), 16), dtype=torch.uint8)

    def extract_z_leavers(self, lo, hi, cap, rebase_lo, rebase_hi):
        n = self.s.np
        iz = self.s.p["cell"][:n] // self.plane
        out = []
        rb = self.real.itemsize
        for buf, sel, rebase in ((lo, iz == 0, rebase_lo), (hi, iz == self.nz + 1, rebase_hi)):
            idx = np.nonzero(sel)[0]
            m = len(idx)
            assert m <= cap
            b = buf.numpy()
            for j, k in enumerate(PARTICLE_NAMES[:7]):
                b[j * cap * rb: j * cap * rb + m * rb] = self.s.p[k][idx].view(np.uint8)
            b[7 * cap * rb: 7 * cap * rb + m * 4] = (self.s.p["cell"][idx] + rebase).astype(np.int32).view(np.uint8)
            out.append(m)
        keep = np.nonzero((iz != 0) & (iz != self.nz + 1))[0]
        for k in PARTICLE_NAMES:
            self.s.p[k][:len(keep)] = self.s.p[k][keep]
        self.s.np = len(keep)
        return tuple(out)

    def extract_async(self, lo, hi, cap, counts, rebase_lo, rebase_hi):
        n_lo, n_hi = self.extract_z_leavers(lo, hi, cap, rebase_lo, rebase_hi)
        counts[0] = n_lo
        counts[1] = n_hi

    def append_async(self, buf, cap, count):
        self.append(buf, cap, int(count[0]))

    def append(self, buf, cap, n):
        if n == 0:
            return
        rb = self.real.itemsize
        b = buf.numpy()
        n0 = self.s.np
        assert n0 + n <= self.cap
        for j, k in enumerate(PARTICLE_NAMES[:7]):
            self.s.p[k][n0:n0 + n] = b[j * cap * rb: j * cap * rb + n * rb].view(self.real)
        self.s.p["cell"][n0:n0 + n] = b[7 * cap * rb: 7 * cap * rb + n * 4].view(np.int32)
        self.s.np = n0 + n

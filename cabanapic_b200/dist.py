"""Multi-GPU execution of the hot path: one process per GPU, torch.distributed for the plumbing.

The reference is single-process; its only traces of decomposition are the
commented VPIC neighbour logic (src/move_p.h:327-346) and the unused grid_t (src/grid.h:83-111).
Two modes, selected per problem size:

* ``replicated`` (small grids): every rank holds the whole grid and a slice of the particles;
  the accumulators are summed across ranks (all-reduce) exactly where the reference calls
  ``Kokkos::Experimental::contribute`` (example/example.cpp:248), then every rank runs the
  identical, deterministic field solve.
* ``slab`` (large 3-D grids): the z axis is cut into contiguous slabs -- z is the slowest index
  of VOXEL (src/types.h:195), so every z-plane incl. its x/y ghosts is one contiguous run of
  (nx+2)(ny+2) cells.  The ghost planes z=0 / z=nzl+1 of a slab hold the neighbour's data instead
  of the periodic image.  Per step a rank exchanges with its -z/+z neighbours (periodic ring):
    1. after push: the accumulator rows deposited into its ghost planes (added into the
       neighbour's planes nzl / 1) and the particles that ended in a ghost plane;
    2. inside advance_e: the J planes of the periodic fold (src/fields.h:126-183; the z sweep
       of jfx comes second, that of jfy first, jfz has none) and of the ghost copy (:33-98);
    3. after each advance_b: the cB ghost-copy planes.
  Everything else (x/y periodicity, all stencils) is the single-GPU kernels unchanged.

The runners are written against a small *engine* interface so the exchange choreography can be
tested on CPU with gloo (tests/test_dist.py drives them with a CPU test engine); on GPUs the
engine is ``GpuEngine`` over the C ABI.  The plane packing / adding is torch slicing on views of
the context's device arrays (O(surface) plumbing); all O(particles) and O(cells) work stays in the
CUDA kernels.
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import Consts, Context


# ------------------------------------------------------------------------------ layout
def slab_ranges(nz: int, world: int):
    """Balanced contiguous z ranges: rank r owns interior planes [z0, z0+nzl) (0-based)."""
    base, rem = divmod(nz, world)
    out, z0 = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((z0, n))
        z0 += n
    return out


def choose_mode(nx, ny, nz, world, mode="auto"):
    """auto: slabs when every rank gets at least 2 planes and the grid is big enough for the
    plane exchange to be worth it; otherwise replicate the grid."""
    if mode != "auto":
        return mode
    if world > 1 and nz >= 2 * world and (nx + 2) * (ny + 2) * (nz + 2) >= 1 << 18:
        return "slab"
    return "replicated"


class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


# ------------------------------------------------------------------------------ GPU engine
class GpuEngine:
    """The engine interface over one ``Context`` (C ABI).  ``nz`` is the local slab thickness."""

    def __init__(self, nx, ny, nz, max_particles, real=np.float32, device=0, fp_mode=_lib.FP_STRICT,
                 z_periodic=True, solver=_lib.SOLVER_EM, stream=None):
        self.nx, self.ny, self.nz = nx, ny, nz
        self.real = np.dtype(real)
        self.tdtype = torch.float32 if self.real.itemsize == 4 else torch.float64
        self.device = torch.device("cuda", device)
        self.ctx = Context(nx, ny, nz, 1, max_particles=max_particles, real=real, device=device, fp_mode=fp_mode,
                           enable_sort=True, solver=solver)
        # all work of this rank -- our kernels, torch's plane ops, NCCL -- is ordered on ONE torch stream: the
        # current one, or `stream` (the caller then runs everything under torch.cuda.stream(stream); a side
        # stream is what CUDA-graph capture of whole steps needs)
        self.stream = stream
        self.ctx.set_stream((stream or torch.cuda.current_stream(self.device)).cuda_stream)
        if not z_periodic:
            self.ctx.set_axis_periodic(1, 1, 0)
        self.gx, self.gy, self.gz = nx + 2, ny + 2, nz + 2
        self.plane = self.gx * self.gy
        self.nc = self.plane * self.gz
        ts = "<f4" if self.real.itemsize == 4 else "<f8"
        ptr, _, stride = self.ctx.device_ptr(16)
        self._fields = torch.as_tensor(_DevArray(ptr, (9, stride), ts), device=self.device)
        ptr, _, _ = self.ctx.device_ptr(18)
        self._acc = torch.as_tensor(_DevArray(ptr, (self.gz, self.plane * 12), ts), device=self.device)
        self.push_ms = 0.0
        self.profile_push = False
        self._push_events = []

    # views ---------------------------------------------------------------------------
    def field_planes(self, m):
        """member m as [gz, plane]"""
        return self._fields[m, :self.nc].view(self.gz, self.plane)

    def acc_planes(self):
        return self._acc

    # compute -------------------------------------------------------------------------
    def load_interpolator(self): self.ctx.load_interpolator_array()
    def clear_accumulator(self): self.ctx.clear_accumulator_array()

    def push(self, k):
        self.ctx.push(k)

    def push_reorder(self, k):
        """push + cell ordering of the store in one pass (cpic_push_reorder)"""
        if self.profile_push and self.async_migration:      # no host wait: events now, elapsed times at the end
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            self.ctx.push_reorder(k)
            ev1.record()
            self._push_events.append((ev0, ev1))
        else:
            self.ctx.push_reorder(k)

    def collect_push_ms(self):
        """device time of the pushes recorded since profiling was switched on (synchronises)"""
        if self._push_events:
            torch.cuda.synchronize(self.device)
            self.push_ms += sum(a.elapsed_time(b) for a, b in self._push_events)
            self._push_events = []
        return self.push_ms

    # migration with the counts on the device (cpic_slab_extract_async / cpic_slab_append_async): float
    # reordering push only
    @property
    def async_migration(self):
        return self.real.itemsize == 4

    def extract_async(self, lo, hi, cap, counts, rebase_lo, rebase_hi):
        self.ctx.slab_extract_async(lo.data_ptr(), hi.data_ptr(), cap, counts.data_ptr(), rebase_lo, rebase_hi)

    def append_async(self, buf, cap, count):
        self.ctx.slab_append_async(buf.data_ptr(), cap, count.data_ptr())

    def unload_accumulator(self, k): self.ctx.unload_accumulator_array(k)
    def fold_phase(self, phase): self.ctx.update_ghosts(3 + phase)
    def ghost_copy_local(self, which): self.ctx.update_ghosts(1 if which == "J" else 2)
    def advance_b_stencil(self, px, py, pz): self.ctx.advance_b_stencil(px, py, pz)
    def advance_e_stencil(self, px, py, pz, cj): self.ctx.advance_e_stencil(px, py, pz, cj)
    def advance_b(self, px, py, pz): self.ctx.advance_b(px, py, pz)
    def advance_e(self, px, py, pz, cj): self.ctx.advance_e(px, py, pz, cj)
    def sort(self): self.ctx.sort_particles()
    def energies(self): return self.ctx.energies()

    # particles -----------------------------------------------------------------------
    @property
    def num_particles(self): return self.ctx.num_particles

    def alloc_bytes(self, n):
        return torch.empty(max(int(n), 16), dtype=torch.uint8, device=self.device)

    def extract_z_leavers(self, lo, hi, cap, rebase_lo, rebase_hi):
        n = self.ctx.extract_z_leavers(lo.data_ptr(), hi.data_ptr(), cap, rebase_lo, rebase_hi)
        if self.profile_push and not self._push_events:   # extract synchronises: the push of this step has finished
            self.push_ms += self.ctx.last_ms(0)
        return n

    def append(self, buf, cap, n):
        self.ctx.append_particles_device(buf.data_ptr(), cap, n)

    def sync(self):
        torch.cuda.synchronize(self.device)

    def close(self):
        self.ctx.close()


# ------------------------------------------------------------------------------ exchange helpers
def _ring_exchange(send_up, send_down, recv_from_down, recv_from_up, up, down, rank):
    """send_up -> rank `up`, send_down -> rank `down`; receive the matching tensors.  With two ranks
    both neighbours are the same peer: the op order (sends: up, down; receives: from-down, from-up)
    keeps the pairs matched on in-order transports (NCCL and gloo).  With one rank it is a copy."""
    if up == rank:                                   # single rank: periodic with itself
        recv_from_down.copy_(send_up)
        recv_from_up.copy_(send_down)
        return
    ops = [dist.P2POp(dist.isend, send_up, up), dist.P2POp(dist.isend, send_down, down),
           dist.P2POp(dist.irecv, recv_from_down, down), dist.P2POp(dist.irecv, recv_from_up, up)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def _ring_exchange_many(items, up, down, rank):
    """items: (send_up, send_down, recv_from_down, recv_from_up) tuples, all in ONE batch of point-to-point
    operations (same pairing rule as _ring_exchange, item after item)."""
    if up == rank:
        for su, sd, rd, ru in items:
            rd.copy_(su)
            ru.copy_(sd)
        return
    ops = []
    for su, sd, _, _ in items:
        ops += [dist.P2POp(dist.isend, su, up), dist.P2POp(dist.isend, sd, down)]
    for _, _, rd, ru in items:
        ops += [dist.P2POp(dist.irecv, rd, down), dist.P2POp(dist.irecv, ru, up)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def _pack(buf, cap, n, rbytes):
    """first n particles of a capacity-`cap` SoA send buffer -> contiguous bytes (same layout, capacity n)"""
    parts = [buf[m * cap * rbytes: m * cap * rbytes + n * rbytes] for m in range(7)]
    parts.append(buf[7 * cap * rbytes: 7 * cap * rbytes + n * 4])
    return torch.cat(parts) if n > 0 else buf[:0]


# ------------------------------------------------------------------------------ slab runner
class SlabStepper:
    """One z-slab of the global box on this rank; ``step()`` is one reference time step
    (example/example.cpp:221-266) with the neighbour exchanges woven in."""

    def __init__(self, engine, k: Consts, rank, world, nzl_down, nzl_up, send_capacity):
        self.e, self.k, self.rank, self.world = engine, k, rank, world
        self.up, self.down = (rank + 1) % world, (rank - 1) % world
        e = engine
        self.rbytes = e.real.itemsize
        self.cap = int(send_capacity)
        nb = self.cap * (7 * self.rbytes + 4)
        self.send_lo, self.send_hi = e.alloc_bytes(nb), e.alloc_bytes(nb)
        # cell re-basing into the receiver's numbering: my ghost plane nzl+1 is the upper neighbour's
        # plane 1; my ghost plane 0 is the lower neighbour's plane nzl_down
        self.rebase_hi = -e.nz * e.plane
        self.rebase_lo = nzl_down * e.plane
        self._migrated = [0, 0]         # particles sent down / up over the runner's life
        self._last_migration = (0, 0)
        # device-counted migration (no host synchronisation inside a step): used by fused steps when the engine has it
        self.async_ok = bool(getattr(e, "async_migration", False))
        if self.async_ok:
            dev = self.send_lo.device
            self.recv_dn, self.recv_up = e.alloc_bytes(nb), e.alloc_bytes(nb)
            self.cnt_send = torch.zeros(2, dtype=torch.int64, device=dev)     # (n_lo, n_hi) of the last extraction
            self.cnt_recv = torch.zeros(2, dtype=torch.int64, device=dev)     # (from the lower, from the upper neighbour)
            self.cnt_total = torch.zeros(2, dtype=torch.int64, device=dev)
            self._async_used = False
        self._last_async = False
        R = engine.real.type
        self.half = (float(R(0.5) * R(k.px)), float(R(0.5) * R(k.py)), float(R(0.5) * R(k.pz)))
        self.nsteps = 0

    @property
    def migrated(self):
        if self.async_ok and self._async_used:
            t = self.cnt_total.tolist()
            return [self._migrated[0] + int(t[0]), self._migrated[1] + int(t[1])]
        return self._migrated

    @property
    def last_migration(self):
        if self.async_ok and self._async_used and self._last_async:
            return tuple(int(v) for v in self.cnt_send.tolist())
        return self._last_migration

    # -- pieces -----------------------------------------------------------------------
    def _exchange_planes(self, members, src_up, src_down, dst_from_down, dst_from_up, add=False):
        """For every field member m: send plane src_up to the upper and src_down to the lower
        neighbour, receive into planes dst_from_down / dst_from_up (assign or add)."""
        e = self.e
        P = [e.field_planes(m) for m in members]
        s_up = torch.stack([p[src_up] for p in P])
        s_dn = torch.stack([p[src_down] for p in P])
        r_dn, r_up = torch.empty_like(s_up), torch.empty_like(s_dn)
        _ring_exchange(s_up, s_dn, r_dn, r_up, self.up, self.down, self.rank)
        for i, p in enumerate(P):
            if add:
                p[dst_from_down] += r_dn[i]
                p[dst_from_up] += r_up[i]
            else:
                p[dst_from_down] = r_dn[i]
                p[dst_from_up] = r_up[i]

    def _exchange_accumulators(self):
        e = self.e
        A = e.acc_planes()
        nz = e.nz
        s_up, s_dn = A[nz + 1].clone(), A[0].clone()
        r_dn, r_up = torch.empty_like(s_up), torch.empty_like(s_dn)
        _ring_exchange(s_up, s_dn, r_dn, r_up, self.up, self.down, self.rank)
        A[1] += r_dn            # the lower neighbour's high ghost plane is my plane 1
        A[nz] += r_up           # the upper neighbour's low ghost plane is my plane nz
        A[0].zero_()
        A[nz + 1].zero_()

    def _migrate(self):
        e = self.e
        n_lo, n_hi = e.extract_z_leavers(self.send_lo, self.send_hi, self.cap, self.rebase_lo, self.rebase_hi)
        self._last_migration = (n_lo, n_hi)
        self._last_async = False
        self._migrated[0] += n_lo
        self._migrated[1] += n_hi
        dev = self.send_lo.device
        cnt_s_up = torch.tensor([n_hi], dtype=torch.int64, device=dev)
        cnt_s_dn = torch.tensor([n_lo], dtype=torch.int64, device=dev)
        cnt_r_dn, cnt_r_up = torch.zeros_like(cnt_s_up), torch.zeros_like(cnt_s_dn)
        _ring_exchange(cnt_s_up, cnt_s_dn, cnt_r_dn, cnt_r_up, self.up, self.down, self.rank)
        m_dn, m_up = int(cnt_r_dn.item()), int(cnt_r_up.item())
        per = 7 * self.rbytes + 4
        p_up = _pack(self.send_hi, self.cap, n_hi, self.rbytes)
        p_dn = _pack(self.send_lo, self.cap, n_lo, self.rbytes)
        r_dn, r_up = e.alloc_bytes(m_dn * per), e.alloc_bytes(m_up * per)
        if self.up == self.rank:
            r_dn[:p_up.numel()].copy_(p_up)
            r_up[:p_dn.numel()].copy_(p_dn)
        else:
            ops = []
            if n_hi: ops.append(dist.P2POp(dist.isend, p_up, self.up))
            if n_lo: ops.append(dist.P2POp(dist.isend, p_dn, self.down))
            if m_dn: ops.append(dist.P2POp(dist.irecv, r_dn[:m_dn * per], self.down))
            if m_up: ops.append(dist.P2POp(dist.irecv, r_up[:m_up * per], self.up))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        e.append(r_dn, m_dn, m_dn)
        e.append(r_up, m_up, m_up)

    def _exchange_after_push_async(self):
        """Ghost accumulator planes + migrating particles in ONE batch of sends/receives, every count on the
        device: nothing here waits for the push on the host, so the host enqueues the rest of the step (and
        the next steps) while the push is still running.  Particle payloads travel as whole capacity-sized
        buffers with their counts alongside."""
        e = self.e
        A = e.acc_planes()
        nz = e.nz
        e.extract_async(self.send_lo, self.send_hi, self.cap, self.cnt_send, self.rebase_lo, self.rebase_hi)
        a_dn, a_up = torch.empty_like(A[0]), torch.empty_like(A[0])
        _ring_exchange_many([(A[nz + 1], A[0], a_dn, a_up),
                             (self.cnt_send[1:2], self.cnt_send[0:1], self.cnt_recv[0:1], self.cnt_recv[1:2]),
                             (self.send_hi, self.send_lo, self.recv_dn, self.recv_up)], self.up, self.down, self.rank)
        A[1] += a_dn            # the lower neighbour's high ghost plane is my plane 1
        A[nz] += a_up           # the upper neighbour's low ghost plane is my plane nz
        A[0].zero_()
        A[nz + 1].zero_()
        e.append_async(self.recv_dn, self.cap, self.cnt_recv[0:1])
        e.append_async(self.recv_up, self.cap, self.cnt_recv[1:2])
        self.cnt_total += self.cnt_send
        self._async_used = True
        self._last_async = True

    def _advance_b(self):
        e = self.e
        e.advance_b_stencil(*self.half)                  # src/fields.h:692-717
        e.ghost_copy_local("B")                          # :718, x and y faces
        nz = e.nz
        self._exchange_planes((3, 4, 5), nz, 1, 0, nz + 1)   # z faces: my top plane is the upper nbr's ghost 0

    def _advance_e(self):
        e, k = self.e, self.k
        nz = e.nz
        gx, gy, nx, ny = e.gx, e.gy, e.nx, e.ny
        # --- periodic fold of J (src/fields.h:126-183), z sweeps replaced by the plane exchange
        e.fold_phase(0)                                  # jfx y-fold, (jfy z-fold skipped), jfz x-fold
        jx, jy = e.field_planes(6), e.field_planes(7)
        s_up = torch.stack([jx[nz + 1], jy[nz + 1]])     # upper ghost plane -> the upper nbr's plane 1
        r_dn = torch.empty_like(s_up)
        if self.up == self.rank:
            r_dn.copy_(s_up)
        else:
            ops = [dist.P2POp(dist.isend, s_up, self.up), dist.P2POp(dist.irecv, r_dn, self.down)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        rx, ry = r_dn[0].view(gy, gx), r_dn[1].view(gy, gx)
        jy[1].view(gy, gx)[1:ny + 1, 1:nx + 2] += ry[1:ny + 1, 1:nx + 2]     # jfy's first sweep (:146-151)
        e.fold_phase(1)                                  # (jfx z-fold skipped), jfy x-fold, jfz y-fold
        jx[1].view(gy, gx)[1:ny + 2, 1:nx + 1] += rx[1:ny + 2, 1:nx + 1]     # jfx's second sweep (:136-141)
        # --- ghost copy of J (:643): x, y locally, z planes from the neighbours
        e.ghost_copy_local("J")
        self._exchange_planes((6, 7, 8), nz, 1, 0, nz + 1)
        e.advance_e_stencil(k.px, k.py, k.pz, k.dt_eps0)  # :646-664

    # -- one step -----------------------------------------------------------------------
    def step(self, sort=False, fused=False):
        e, k = self.e, self.k
        t = self._tick
        t(None)
        if sort:
            e.sort()
        e.load_interpolator()
        e.clear_accumulator()
        t("sort+interp")
        if fused:
            e.push_reorder(k)
        else:
            e.push(k)
        t("push")
        if fused and self.async_ok:
            self._exchange_after_push_async()
            t("acc exchange + migrate (async)")
        else:
            self._exchange_accumulators()
            t("acc exchange")
            self._migrate()
            t("migrate")
        e.unload_accumulator(k)
        self._advance_b()
        self._advance_e()
        self._advance_b()
        t("fields+exchange")
        self.nsteps += 1

    # developer phase profile (CPIC_SLAB_PROFILE=1): host wall time per phase with a device sync at every
    # phase boundary -- it serialises the step, so the numbers explain a step, they do not time it
    _prof = None

    def _tick(self, name):
        if self._prof is None:
            import os
            SlabStepper._prof = {} if os.environ.get("CPIC_SLAB_PROFILE") else False
        if self._prof is False:
            return
        torch.cuda.synchronize()
        now = time.perf_counter()
        if name is not None:
            self._prof[name] = self._prof.get(name, 0.0) + (now - self._t_last) * 1e3
        self._t_last = now

    def profile_report(self):
        return dict(self._prof) if self._prof else None

    def energies(self):
        eng, b = self.e.energies()
        t = torch.tensor([eng, b], dtype=torch.float64)
        if self.world > 1:
            t = t.to(self.send_lo.device)
            dist.all_reduce(t)
        return float(t[0]), float(t[1])


# ------------------------------------------------------------------------------ replicated runner
class ReplicatedStepper:
    """Whole grid on every rank, particles split by index; accumulators all-reduced."""

    def __init__(self, engine, k: Consts, rank, world):
        self.e, self.k, self.rank, self.world = engine, k, rank, world
        R = engine.real.type
        self.half = (float(R(0.5) * R(k.px)), float(R(0.5) * R(k.py)), float(R(0.5) * R(k.pz)))

    def step(self, sort=False, fused=False):
        e, k = self.e, self.k
        if sort:
            e.sort()
        e.load_interpolator()
        e.clear_accumulator()
        if fused:
            e.push_reorder(k)
        else:
            e.push(k)
        if self.world > 1:
            dist.all_reduce(e.acc_planes())            # contribute(), example/example.cpp:248
        e.unload_accumulator(k)
        e.advance_b(*self.half)
        e.advance_e(k.px, k.py, k.pz, k.dt_eps0)
        e.advance_b(*self.half)

    def energies(self):
        return self.e.energies()                       # fields are replicated: every rank has the total


# ------------------------------------------------------------------------------ bench runners
class _BenchRunner:
    """Adapter with the interface bench.py drives (see SingleGpu there)."""

    def __init__(self, d, k, we, rank, world, local, fp_mode):
        self.d, self.k, self.we, self.rank, self.world, self.local, self.fp = d, k, we, rank, world, local, fp_mode
        self.l0 = 0
        self.t_ms = 0.0

    # Multi-GPU steps are short (4 ms at 8 GPUs) and made of ~120 host calls (torch plane ops, NCCL groups, our
    # launches): with the slab exchange free of host synchronisation (cpic_slab_*_async) a PAIR of steps -- the
    # particle double buffer is back where it started after two -- is captured into one CUDA graph and replayed.
    graph = None
    used_graph = False
    graph_launches = 0
    host_ms = 0.0
    last_n = 0
    _stream = None

    def _on_stream(self):
        import contextlib
        return torch.cuda.stream(self._stream) if self._stream is not None else contextlib.nullcontext()

    def prepare_timed(self, sort_interval):
        """called once after the warm-up steps: capture two fused steps (falls back to eager launches on any error)"""
        import os
        st = self.stepper
        if (self.graph is not None or self._stream is None or sort_interval >= 0 or os.environ.get("CPIC_NO_GRAPH")
                or not getattr(st, "async_ok", False) or not getattr(st, "_async_used", False)):
            return
        keep = self.eng.profile_push
        self.eng.profile_push = False
        try:
            torch.cuda.synchronize()
            l0 = self.eng.ctx.launch_count
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                st.step(fused=True)
                st.step(fused=True)
            st.nsteps -= 2                       # captured, not executed
            self.graph_launches = self.eng.ctx.launch_count - l0
            self.graph = g
            self.used_graph = True
        except Exception as ex:                  # pragma: no cover - depends on the NCCL / driver at hand
            import sys
            print(f"[rank {self.rank}] CUDA-graph capture of the slab step failed, running eagerly: {ex}", file=sys.stderr, flush=True)
            self.graph = None
            torch.cuda.synchronize()
        self.eng.profile_push = keep

    def step(self, n, sort_interval):
        self.l0 = self.eng.ctx.launch_count
        self.last_n = n
        with self._on_stream():
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ev0.record()
            s = 0
            if self.graph is not None and sort_interval < 0:
                while s + 2 <= n:
                    self.graph.replay()
                    s += 2
                self.stepper.nsteps += s
                self.l0 -= self.graph_launches * (s // 2)
            keep = self.eng.profile_push
            if self.graph is not None:
                self.eng.profile_push = False          # (push times come from an eager pass, profile_result)
            for s in range(s, n):
                self.stepper.step(sort=sort_interval > 0 and s % sort_interval == 0, fused=sort_interval < 0)
            self.eng.profile_push = keep
            ev1.record()
            self.host_ms = (time.perf_counter() - t0) * 1e3
            torch.cuda.synchronize()
            self.t_ms = ev0.elapsed_time(ev1)

    def profile(self, on):
        self.eng.profile_push = bool(on)
        self.eng.push_ms = 0.0

    def profile_result(self):
        if self.graph is not None and self.eng.profile_push:
            # the timed region replayed a graph: time the same push kernel over a few eager steps right after it
            with self._on_stream():
                self.eng.push_ms = 0.0
                for _ in range(4):
                    self.stepper.step(fused=True)
                ms = self.eng.collect_push_ms() / 4
            return {"push_ms": ms * max(self.last_n, 1), "steps": 0, "push_timing": "4 eager steps after the graph-replayed timed region"}
        with self._on_stream():
            return {"push_ms": self.eng.collect_push_ms(), "steps": 0}

    def device_ms(self):
        return self.t_ms

    def launches_in_timed_region(self):
        return self.eng.ctx.launch_count - self.l0

    def local_particles(self):
        return self.eng.num_particles

    def e2e(self, steps, sort_interval):
        """Per step: H2D of this rank's particles + fields from pinned host memory, one step with
        all exchanges, D2H of particles + fields (every rank, concurrently)."""
        import ctypes as C
        import psutil
        c = self.eng.ctx
        n = c.num_particles
        nbytes = n * 32 + 9 * c.nc * 4
        if psutil.virtual_memory().available < 1.6 * nbytes * self.world:
            return None
        names = "dx dy dz ux uy uz w".split()
        host = {m: torch.empty(n, dtype=torch.float32, pin_memory=True).numpy() for m in names}
        host["cell"] = torch.empty(n, dtype=torch.int32, pin_memory=True).numpy()
        cap2 = int(n * 1.05) + 1024                       # the count drifts by the net migration
        back = {m: torch.empty(cap2, dtype=torch.float32, pin_memory=True).numpy() for m in names}
        back["cell"] = torch.empty(cap2, dtype=torch.int32, pin_memory=True).numpy()
        hf = torch.empty((9, c.nc), dtype=torch.float32, pin_memory=True).numpy()
        L = c.L
        up = [host[m].ctypes.data_as(C.c_void_p) for m in names] + [host["cell"].ctypes.data_as(C.c_void_p)]
        dn = [back[m].ctypes.data_as(C.c_void_p) for m in names] + [back["cell"].ctypes.data_as(C.c_void_p)]
        fptr = (C.c_void_p * 9)(*[hf[m].ctypes.data for m in range(9)])
        got = C.c_int64()
        c._ck(L.cpic_download_particles(c.h, *up, n, C.byref(got)))
        c._ck(L.cpic_download_fields(c.h, fptr))
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with self._on_stream():
            for s in range(steps):
                c._ck(L.cpic_upload_particles(c.h, *up, n))
                c._ck(L.cpic_upload_fields(c.h, fptr))
                self.stepper.step(sort=sort_interval > 0, fused=sort_interval < 0)
                c._ck(L.cpic_download_particles(c.h, *dn, cap2, C.byref(got)))
                c._ck(L.cpic_download_fields(c.h, fptr))
            if self.world > 1:
                dist.barrier()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        return {"value": self.d.num_particles * steps / sec, "unit": "particle-steps/s",
                "h2d_bytes_per_step": nbytes * self.world, "d2h_bytes_per_step": nbytes * self.world,
                "steps": steps, "seconds": sec,
                "what": "per step and rank: H2D local particles+fields from pinned host, one step incl. NCCL "
                        "exchanges, D2H local particles+fields; bytes summed over ranks"}

    def close(self):
        if self.graph is not None:           # the graph references NCCL work: release it before the communicator goes
            torch.cuda.synchronize()
            self.graph = None
            import gc
            gc.collect()
            torch.cuda.synchronize()
        rep = getattr(self.stepper, "profile_report", lambda: None)()
        if rep and self.rank == 0:
            import sys
            n = max(getattr(self.stepper, "nsteps", 1), 1)
            print("slab phase profile, ms/step (serialised): " + ", ".join(f"{k} {v / n:.3f}" for k, v in rep.items()),
                  file=sys.stderr, flush=True)
        self.eng.close()


class SlabBench(_BenchRunner):
    open_z = False      # tests: keep the z exchange (with itself) even on one rank

    def setup(self):
        d = self.d
        z0, nzl = slab_ranges(d.nz, self.world)[self.rank]
        ranges = slab_ranges(d.nz, self.world)
        per_plane = d.nx * d.ny * d.nppc
        n_local = per_plane * nzl
        cap = int(n_local * 1.10) + 4096
        self._stream = torch.cuda.Stream(device=self.local)
        with self._on_stream():
            self.eng = GpuEngine(d.nx, d.ny, nzl, cap, real=d.real, device=self.local, fp_mode=self.fp,
                                 z_periodic=self.world == 1 and not self.open_z, stream=self._stream)
            self.eng.ctx.init_uniform_plasma(z0 * per_plane, n_local, d.nx, d.ny, d.nz, d.nppc, z0=z0, weight=self.we)
        # ~2.3 % of a plane's particles cross a z face per step in this plasma (vth = 0.1 c, dt = 0.99 Courant);
        # the exchange buffers travel whole, so the capacity is kept near 2x that (an overflow is reported)
        send_cap = max(4096, int(per_plane * 0.05))
        nzl_down = ranges[(self.rank - 1) % self.world][1]
        nzl_up = ranges[(self.rank + 1) % self.world][1]
        with self._on_stream():
            self.stepper = SlabStepper(self.eng, self.k, self.rank, self.world, nzl_down, nzl_up, send_cap)
        self.eng.sync()

    def describe(self):
        how = "two steps per CUDA-graph replay" if self.used_graph else "eager launches"
        return f"{self.world} z-slabs (NCCL ghost-plane exchange + particle migration, {how})"


class ReplicatedBench(_BenchRunner):
    def setup(self):
        d = self.d
        n = d.num_particles
        first = n * self.rank // self.world
        last = n * (self.rank + 1) // self.world
        self.eng = GpuEngine(d.nx, d.ny, d.nz, last - first, real=d.real, device=self.local, fp_mode=self.fp)
        self.eng.ctx.init_uniform_plasma(first, last - first, d.nx, d.ny, d.nz, d.nppc, weight=self.we)
        self.stepper = ReplicatedStepper(self.eng, self.k, self.rank, self.world)
        self.eng.profile_hook = True
        self.eng.sync()

    def step(self, n, sort_interval):
        super().step(n, sort_interval)
        if self.eng.profile_push:                         # no per-step sync here: time the last push only
            self.eng.push_ms += self.eng.ctx.last_ms(0) * n

    def describe(self):
        return f"grid replicated on {self.world} GPUs, particles split, NCCL all-reduce of the accumulator"


class NativeBench:
    """bench.py runner over the NATIVE multi-GPU layer (include/cabanapic_b200_mgpu.h): the whole step -- kernels, plane
    exchange, particle migration, NCCL calls, CUDA-graph replay -- is host C++ inside the library; torch.distributed
    only hands the 128-byte NCCL id to the ranks and reduces the timings."""

    def __init__(self, d, k, we, rank, world, local, fp_mode, mode):
        self.d, self.k, self.we, self.rank, self.world, self.local, self.fp, self.mode = d, k, we, rank, world, local, fp_mode, mode
        self.l0 = 0
        self.t_ms = 0.0
        self.host_ms = 0.0
        self.used_graph = False
        self.use_graph = True
        self._profile = False
        self._push_ms = 0.0

    def setup(self):
        import os
        d = self.d
        dev = torch.device("cuda", self.local)
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if self.rank == 0:
            idt.copy_(torch.frombuffer(bytearray(_lib.Mgpu.unique_id()), dtype=torch.uint8))
        if self.world > 1:
            dist.broadcast(idt, 0)
        uid = bytes(idt.cpu().numpy().tobytes())
        per_plane = d.nx * d.ny * d.nppc
        if self.mode == "slab":
            nzl = slab_ranges(d.nz, self.world)[self.rank][1]
            cap = int(per_plane * nzl * 1.10) + 4096
            send_cap = max(4096, int(per_plane * 0.05))
            mode = _lib.MGPU_SLAB
        else:
            n = d.num_particles
            cap = n * (self.rank + 1) // self.world - n * self.rank // self.world
            send_cap, mode = 0, _lib.MGPU_REPLICATED
        self.use_graph = not os.environ.get("CPIC_NO_GRAPH")
        self.m = _lib.Mgpu(d.nx, d.ny, d.nz, self.rank, self.world, uid, mode=mode, max_particles=cap, real=d.real,
                           device=self.local, fp_mode=self.fp, send_capacity=send_cap)
        self.transport = self.m.transport
        self.m.init_uniform_plasma(d.nppc, weight=self.we)
        self.m.sync()

    def step(self, n, sort_interval):
        c = self.m.ctx
        self.l0 = c.launch_count
        t0 = time.perf_counter()
        self.m.step(self.k, n, sort_interval, use_graph=self.use_graph)
        self.host_ms = (time.perf_counter() - t0) * 1e3
        self.m.sync()
        self.used_graph = self.used_graph or self.m.used_graph
        self.t_ms = c.last_ms(3)
        if self._profile:
            self._push_ms += c.last_ms(0) * n          # no per-step sync: the last push of the call stands for all of them

    def prepare_timed(self, sort_interval):
        """after the warm-up steps: capture the pair-of-steps graph now (nothing executes), so that the timed call only
        replays it"""
        if self.use_graph and sort_interval < 0 and self.mode == "slab" and self.world > 1:
            self.m.prepare_graph(self.k)

    def profile(self, on):
        self._profile = bool(on)
        self._push_ms = 0.0

    def profile_result(self):
        return {"push_ms": self._push_ms, "steps": 0, "push_timing": "device time of the last push of the timed call x steps"}

    def device_ms(self):
        return self.t_ms

    def launches_in_timed_region(self):
        return self.m.ctx.launch_count - self.l0

    def local_particles(self):
        return self.m.ctx.num_particles

    def digest(self):
        return self.m.state_digest()

    def describe(self):
        how = "two steps per CUDA-graph replay" if self.used_graph else "eager launches"
        if self.mode == "slab":
            via = {"peer-memory": "peer-memory (NVLink stores into the neighbours' mailboxes / ghost planes, CUDA IPC)",
                   "nccl": "NCCL send/recv"}.get(self.transport, self.transport)
            return (f"{self.world} z-slabs, native C++ stepper (cpic_mgpu_step): ghost-plane exchange + device-counted "
                    f"particle migration over {via}, {how}")
        return f"grid replicated on {self.world} GPUs, particles split, ncclAllReduce of the accumulator (cpic_mgpu_step)"

    def e2e(self, steps, sort_interval):
        """Per step and rank: ONE cpic_mgpu_step_host call on pinned host arrays -- this rank's slab streams through its
        GPU in chunks (H2D / in-place push / D2H overlapped), the ghost-plane and leaver exchanges and the field advance
        follow, and the host copy is patched where the migration changed it.  (Replicated mode: upload -> step ->
        download.)"""
        import ctypes as C
        import psutil
        c = self.m.ctx
        n = c.num_particles
        nbytes = n * 32 + 9 * c.nc * 4
        ok = psutil.virtual_memory().available >= 2.3 * nbytes * self.world
        if self.world > 1:                                # every rank takes the same decision
            flag = torch.tensor([1 if ok else 0], device="cuda", dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = bool(flag.item())
        if not ok:
            return None
        names = "dx dy dz ux uy uz w".split()
        cap2 = int(n * 1.05) + 1024                       # the count drifts by the net migration
        # the host arrays of a rank belong on the NUMA node its GPU hangs off (what numactl / the MPI launcher of a real
        # run would arrange): pin this process to the GPU's CPU set while the pinned buffers are allocated and touched
        old_aff = _bind_to_gpu_cpus(self.local) if self.world > 1 else None
        host = {m: torch.zeros(cap2, dtype=torch.float32, pin_memory=True).numpy() for m in names}
        host["cell"] = torch.zeros(cap2, dtype=torch.int32, pin_memory=True).numpy()
        back = {m: torch.zeros(cap2, dtype=torch.float32, pin_memory=True).numpy() for m in names}
        back["cell"] = torch.zeros(cap2, dtype=torch.int32, pin_memory=True).numpy()
        hf = torch.zeros((9, c.nc), dtype=torch.float32, pin_memory=True).numpy()
        hf2 = torch.zeros((9, c.nc), dtype=torch.float32, pin_memory=True).numpy()
        L = c.L
        up = [host[m].ctypes.data_as(C.c_void_p) for m in names] + [host["cell"].ctypes.data_as(C.c_void_p)]
        fptr = (C.c_void_p * 9)(*[hf[m].ctypes.data for m in range(9)])
        got = C.c_int64()
        c._ck(L.cpic_download_particles(c.h, *up, cap2, C.byref(got)))
        c._ck(L.cpic_download_fields(c.h, fptr))
        cnt = int(got.value)
        streamed = self.mode == "slab"
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        moved = 0
        for s in range(steps):
            moved += cnt
            if streamed:
                cnt = self.m.step_host(self.k, host, back, cnt, hf, hf2)
                host, back, hf, hf2 = back, host, hf2, hf
            else:
                up = [host[m].ctypes.data_as(C.c_void_p) for m in names] + [host["cell"].ctypes.data_as(C.c_void_p)]
                fptr = (C.c_void_p * 9)(*[hf[m].ctypes.data for m in range(9)])
                c._ck(L.cpic_upload_particles(c.h, *up, cnt))
                c._ck(L.cpic_upload_fields(c.h, fptr))
                self.m.step(self.k, 1, sort_interval, use_graph=False)
                c._ck(L.cpic_download_particles(c.h, *up, cap2, C.byref(got)))
                c._ck(L.cpic_download_fields(c.h, fptr))
                cnt = int(got.value)
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        if old_aff is not None:
            import os
            os.sched_setaffinity(0, old_aff)
        what = ("per step and rank: ONE cpic_mgpu_step_host call on pinned host arrays (this slab's 8 particle members + 9 field "
                "components in, the same out): chunked H2D / in-place push / D2H pipeline, peer-memory exchanges, field "
                "advance, host copy patched for the migration; bytes summed over ranks") if streamed else (
                "per step and rank: H2D local particles+fields from pinned host, one cpic_mgpu_step, D2H local particles+fields; "
                "bytes summed over ranks")
        return {"value": self.d.num_particles * steps / sec, "unit": "particle-steps/s",
                "h2d_bytes_per_step": nbytes * self.world, "d2h_bytes_per_step": nbytes * self.world,
                "steps": steps, "seconds": sec, "what": what, "host_buffers_numa_local": old_aff is not None}

    def close(self):
        self.m.close()


def _bind_to_gpu_cpus(local):
    """Restrict this process to the CPUs NVML reports as local to GPU `local` (its NUMA node); returns the previous
    affinity set, or None when that is not possible."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        old = os.sched_getaffinity(0)
        cpus &= old
        if not cpus or cpus == old:
            return None
        os.sched_setaffinity(0, cpus)
        return old
    except Exception:
        return None


def make_runner(d, k, we, rank, world, local, mode="auto", fp_mode=_lib.FP_STRICT, native=True):
    mode = choose_mode(d.nx, d.ny, d.nz, world, mode)
    if native and np.dtype(d.real).itemsize == 4:
        return NativeBench(d, k, we, rank, world, local, fp_mode, mode)
    cls = SlabBench if mode == "slab" else ReplicatedBench
    return cls(d, k, we, rank, world, local, fp_mode)

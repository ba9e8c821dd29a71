#!/usr/bin/env python
"""Benchmark of the CabanaPIC particle hot path on B200 (contract: see the build prompt).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N ...            # reference CPU arm (rank 0 only)

metric    particle-steps/s of the whole time step (interpolator load + push/move/deposit [+ the cell
          ordering folded into the push] + accumulator unload + Yee field advance) on the synthetic uniform thermal plasma of
          BASELINE.json configs[4] / SURVEY.md §8(d): 256^3 cells x 64 particles/cell = 2^30
          particles, float, periodic, vth = 0.1 c, dt = 0.99 Courant.
value     device-resident throughput (state already in HBM), CUDA events on the context's stream.
e2e       the same metric through the C-ABI with HOST buffers: every e2e step is one cpic_step_host
          call that takes the whole particle + field state from pinned host memory and returns
          the advanced particles, fields and energies there (what a caller holding its state in
          host arrays pays); the particles stream through the device in chunks, H2D / push / D2H
          overlapped, so the step is bounded by one PCIe direction.
roofline  the push kernel: 56 algorithmic bytes per particle-step x particles per
          launch / the kernel's mean duration (CUDA events around every push launch), against
          the measured HBM copy bandwidth in MEASURED_PEAKS.json.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/s (whole step: push+move+deposit+field advance)"      # identical in both arms
BYTES_PER_PARTICLE_STEP = 56.0       # read 8 members (32 B) + write dx,dy,dz,ux,uy,uz (24 B), float
FALLBACK_HBM_GBS = 6650.0            # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, nargs=3, default=None, help="override the workload grid (development)")
    ap.add_argument("--nppc", type=int, default=64)
    ap.add_argument("--sort-interval", type=int, default=None)
    ap.add_argument("--fp", default="strict", choices=["strict", "contract"])
    ap.add_argument("--mode", default="auto", choices=["auto", "slab", "replicated"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-extras", action="store_true", help="skip the C1/C2/C3 side lines (N=1 only)")
    ap.add_argument("--python-stepper", action="store_true", help="N>1: the round-1 torch stepper instead of cpic_mgpu_step")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        pw = []
        for r in self.rows:
            try:
                pw.append(float(r[3]))
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_mhz_min": min(sm) if sm else None, "power_w_max": max(pw) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ workload
def workload(args, free_bytes=None):
    grid = tuple(args.grid) if args.grid else (256, 256, 256)
    nppc = args.nppc
    name = f"uniform thermal plasma {grid[0]}x{grid[1]}x{grid[2]} cells x {nppc} ppc (BASELINE configs[4])"
    if free_bytes is not None and not args.grid:
        need = grid[0] * grid[1] * grid[2] * nppc * 64 * 1.08 + 4e9      # two particle buffers + grid arrays
        while need > free_bytes and grid[2] > 8:
            grid = (grid[0], grid[1], grid[2] // 2)
            need = grid[0] * grid[1] * grid[2] * nppc * 64 * 1.08 + 4e9
            name = f"uniform thermal plasma {grid[0]}x{grid[1]}x{grid[2]} cells x {nppc} ppc (z halved to fit HBM)"
    return grid, nppc, name


class stdout_to_stderr:
    """The reference's sources print their deck banners to stdout; the bench's stdout carries ONE JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


# ------------------------------------------------------------------------------ CPU arm
def cpu_sample_grid(grid, nppc):
    """C5 scaled to host RAM -- 128^3 cells x 64 ppc when the host has the memory (state + the
    reference's own copy + one accumulator per thread, ~14 GB), else a z-thin slab of the same plasma."""
    import psutil
    nx, ny, nz = grid
    avail = psutil.virtual_memory().available
    cand = (min(nx, 128), min(ny, 128), min(nz, 128))
    n = cand[0] * cand[1] * cand[2] * nppc
    threads = len(os.sched_getaffinity(0))
    need = n * 32 * 2.3 + (cand[0] + 2) * (cand[1] + 2) * (cand[2] + 2) * (48 * threads + 200)
    if need < 0.5 * avail and threads >= 8:
        return cand
    return (nx, ny, min(nz, 4))


def tiled_plasma(d, we, base=1 << 22):
    """The synthetic plasma for a CPU timing sample: the first `base` Philox particles of the GPU generator (numpy, ~1 us
    each) tiled over the box -- same distributions, same cell-by-cell order, different cells; generating all of them in
    numpy would take minutes."""
    from cabanapic_b200 import decks
    n = d.num_particles
    blk = decks.uniform_plasma_chunk(d, we, 0, min(base, n))
    p = {}
    reps = (n + len(blk["dx"]) - 1) // len(blk["dx"])
    for m in "dx dy dz ux uy uz w".split():
        p[m] = np.tile(blk[m], reps)[:n]
    c = np.arange(n, dtype=np.int64) // d.nppc
    ix, iy, iz = c % d.nx, (c // d.nx) % d.ny, c // (d.nx * d.ny)
    p["cell"] = ((ix + 1) + (d.nx + 2) * ((iy + 1) + (d.ny + 2) * (iz + 1))).astype(np.int32)
    return p


def cpu_reference_run(sample_grid, nppc, steps, warmup, min_seconds=0.0, max_steps=None, one_thread=False):
    """Time the reference's own CPU implementation (oracle/_ref: the reference's sources compiled against the
    OpenMP Kokkos/Cabana stand-in; else the scalar C restatement) on a bounded sample of the same plasma."""
    from cabanapic_b200 import decks
    from oracle.api import Consts as OConsts, RefLib, Restatement, State
    nx, ny, nz = sample_grid
    d = decks.uniform_plasma(nx, ny, nz, nppc)
    k, _, we = d.consts()
    ok = OConsts(**k.to_dict())
    p = tiled_plasma(d, we)
    s = State(nx, ny, nz, 1, d.num_particles, "f32")
    for n in p:
        s.p[n][:] = p[n]
    del p
    sample = f"{nx}x{ny}x{nz} cells x {nppc} ppc = {d.num_particles} particles of the same plasma, float, EM"
    if RefLib.available("default", "f32", omp=True):
        # all the host threads this process may use -- torchrun exports OMP_NUM_THREADS=1 to its workers, which would
        # silently turn the reference arm into a single-thread run
        try:
            import ctypes
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(1 if one_thread else len(os.sched_getaffinity(0)))
        except Exception:
            pass
        R = RefLib("default", "f32", omp=True).create(s, solver=0)
        cores, kind = R.num_threads(), "reference"
        run = lambda n: R.run(ok, n)
    else:
        O = Restatement("f32")
        cores, kind = 1, "port"
        run = lambda n: O.step(s, ok, 0, n)
    run(max(1, warmup))
    t0 = time.perf_counter()
    done = 0
    while True:
        run(steps)
        done += steps
        el = time.perf_counter() - t0
        if el >= min_seconds or (max_steps and done >= max_steps):
            break
    return {"value": d.num_particles * done / el, "unit": "particle-steps/s", "cores": cores, "kind": kind,
            "sample": sample + f", {done} steps in {el:.2f} s",
            "what": "the reference's own src/*.cpp + headers compiled where they lie (oracle/Makefile) against an OpenMP "
                    "Kokkos/Cabana stand-in (neither is installable offline); same loop as example/example.cpp:221-266"}, \
        el / done * 1e3, d.num_particles


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    grid, nppc, name = workload(args)
    sg = cpu_sample_grid(grid, nppc)
    # each "step" is one full time step of the bounded sample
    with stdout_to_stderr():
        cb, ms, npart = cpu_reference_run(sg, nppc, args.steps, args.warmup, min_seconds=0.0, max_steps=args.steps)
    line = {"impl": "reference", "metric": METRIC,
            "value": cb["value"], "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "ran": cb["sample"],
                       "note": "CPU arm: every step is one whole time step of a bounded sample of the workload (the "
                               "full 2^30-particle box needs > 100 GB of host state and ~5 s per step); throughput is "
                               "per particle-step, so it compares with the GPU arm's"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import cabanapic_b200 as cp
    from cabanapic_b200 import decks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    free, total = torch.cuda.mem_get_info()
    grid, nppc, name = workload(args, free_bytes=free * world)
    nx, ny, nz = grid
    d = decks.uniform_plasma(nx, ny, nz, nppc)
    k, _, we = d.consts()
    n_total = d.num_particles
    # -1 = CPIC_SORT_FUSED: the push keeps the store in cell order itself (cpic_push_reorder), no sort pass.
    # Measured at C5 with the record store (profiles/r02_bench_n1_records_*.log): fused 34.0, sort every 8 steps 34.8 ms/step
    sort_interval = args.sort_interval if args.sort_interval is not None else -1
    fp = cp.FP_CONTRACT if args.fp == "contract" else cp.FP_STRICT

    if world > 1:
        from cabanapic_b200 import dist as cdist
        runner = cdist.make_runner(d, k, we, rank, world, local, mode=args.mode, fp_mode=fp, native=not args.python_stepper)
    else:
        runner = SingleGpu(d, k, we, local, fp)
    runner.setup()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        runner.step(1, sort_interval)
    if hasattr(runner, "prepare_timed"):
        runner.prepare_timed(sort_interval)      # multi-GPU: capture two steps into a CUDA graph
    barrier()
    runner.profile(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    runner.step(args.steps, sort_interval)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = runner.device_ms()
    clocks = sampler.stop() if rank == 0 else None
    prof = runner.profile_result()
    launches = runner.launches_in_timed_region()
    # parity block: the state the timed region left behind, digested on the device (collective at N > 1)
    digest = runner.digest() if hasattr(runner, "digest") else None
    if world > 1:
        t = torch.tensor([dev_ms, wall_ms, prof["push_ms"], float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, push_ms_max = t[0].item(), t[1].item(), t[2].item()
        prof["push_ms"] = push_ms_max
        launches = int(t[3].item())
    elapsed_ms = max(dev_ms, 0.0) if world == 1 else wall_ms     # multi-GPU: barrier-to-barrier, max over ranks
    value = n_total * args.steps / (elapsed_ms * 1e-3)

    peak, peak_src = measured_peak()
    push_ms_per_launch = prof["push_ms"] / args.steps
    per_launch_particles = runner.local_particles()
    achieved = BYTES_PER_PARTICLE_STEP * per_launch_particles / (push_ms_per_launch * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_push3 (push + move_p + deposit + cell ordering; block-owned cell chunks)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_particle_step": BYTES_PER_PARTICLE_STEP,
                "particles_per_launch": per_launch_particles, "ms_per_launch": push_ms_per_launch,
                "push_only_particle_steps_per_s_per_gpu": per_launch_particles / (push_ms_per_launch * 1e-3),
                "phase_ms_per_step": {kk: v / args.steps for kk, v in prof.items() if kk.endswith("_ms")}}
    tr = os.path.join(ROOT, "profiles", "push_traffic.json")
    if os.path.exists(tr):
        try:
            t = json.load(open(tr))
            roofline["traffic"] = t["dram_bytes_per_particle"] * per_launch_particles
            roofline["traffic_source"] = t.get("source")
        except Exception:
            pass

    e2e = None
    if not args.no_e2e and world == 1:
        e2e = runner.e2e(args.e2e_steps, sort_interval)
    elif world > 1 and not args.no_e2e:
        e2e = runner.e2e(args.e2e_steps, sort_interval)
        if e2e is not None:
            t = torch.tensor([e2e["seconds"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e["value"] = n_total * e2e["steps"] / t[0].item()
    runner.close()

    cpu_base = None
    extras = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        with stdout_to_stderr():
            cpu_base, _, _ = cpu_reference_run(cpu_sample_grid((nx, ny, nz), nppc), nppc, 2, 1, min_seconds=10.0, max_steps=200)
            try:      # BASELINE.md section 4 also asks for a 1-thread figure
                one, _, _ = cpu_reference_run((64, 64, 16), nppc, 1, 1, min_seconds=2.0, max_steps=8, one_thread=True)
                cpu_base["value_1thread"] = one["value"]
                cpu_base["sample_1thread"] = one["sample"]
            except Exception as ex:      # pragma: no cover
                cpu_base["value_1thread"] = None
                cpu_base["note_1thread"] = str(ex)
    if rank == 0 and world == 1 and not args.no_extras and not args.grid:
        with stdout_to_stderr():
            extras = side_lines(local)

    if rank == 0:
        parity = None
        if digest is not None:
            parity = dict(digest)
            parity["steps_taken"] = args.warmup + args.steps
            parity["expected_particles"] = n_total
            parity["expected_weight_sum"] = float(np.float32(we)) * n_total
            parity["ok"] = bool(digest["particles"] == n_total and digest["cells_not_interior"] == 0 and
                                digest["offsets_out_of_range"] == 0 and
                                abs(digest["weight_sum"] - parity["expected_weight_sum"]) <= 1e-6 * parity["expected_weight_sum"])
        line = {"metric": METRIC,
                "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": name, "particles": n_total, "cells": [nx, ny, nz], "ppc": nppc,
                           "sort_interval": sort_interval, "fp_mode": args.fp,
                           "push": "k_push3: CTA-owned cell chunks, TMA-staged interpolators, native/foreigner split, "
                                   "branch-free segmented sum + red.v4, cell ordering folded into the push",
                           "parallelism": runner.describe(),
                           "l2": "inputs (>= 30 GB per GPU) exceed the 126 MB L2; no flush needed"},
                "clocks": clocks, "gpu_launches": launches, "wall_ms_per_step": wall_ms / args.steps,
                "roofline": roofline}
        if getattr(runner, "host_ms", 0.0):
            line["host_enqueue_ms_per_step"] = runner.host_ms / args.steps
        if e2e is not None:
            e2e.pop("seconds", None)
            line["e2e"] = e2e
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if parity is not None:
            line["parity"] = parity
        if extras:
            line["extra"] = extras
        print(json.dumps(line), flush=True)
    if world > 1:
        # never let a teardown problem (e.g. a communicator still referenced by a captured graph) hang the run:
        # the result is printed; give destroy_process_group a few seconds, then leave
        sys.stdout.flush()
        killer = threading.Timer(20.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        try:
            dist.barrier()
            dist.destroy_process_group()
        finally:
            killer.cancel()


class SingleGpu:
    """One context on one GPU holding the whole box."""

    def __init__(self, d, k, we, device, fp):
        self.d, self.k, self.we, self.device, self.fp = d, k, we, device, fp
        self.l0 = 0

    def setup(self):
        import cabanapic_b200 as cp
        d = self.d
        self.c = cp.Context(d.nx, d.ny, d.nz, 1, max_particles=d.num_particles, real=np.float32, device=self.device,
                            fp_mode=self.fp, enable_sort=True)
        self.c.init_uniform_plasma(0, d.num_particles, d.nx, d.ny, d.nz, d.nppc, weight=self.we)
        self.c.upload_fields(d.initial_fields())
        self.c.sync()

    def step(self, n, sort_interval):
        self.l0 = self.c.launch_count
        self.c.step(self.k, n, sort_interval, False)
        self.c.sync()

    def profile(self, on):
        self.c.enable_step_profile(on)

    def profile_result(self):
        return self.c.step_profile()

    def device_ms(self):
        return self.c.last_ms(3)

    def launches_in_timed_region(self):
        return self.c.launch_count - self.l0

    def local_particles(self):
        return self.c.num_particles

    def digest(self):
        return self.c.state_digest()

    def describe(self):
        return "1 GPU, whole domain"

    def e2e(self, steps, sort_interval):
        """Stateless steps through the C ABI with HOST buffers (pinned): upload particles and
        fields, one step, download particles, fields and energies -- every step."""
        import psutil
        import torch
        c, d = self.c, self.d
        n = c.num_particles
        nbytes = n * 32 + 9 * c.nc * 4
        if psutil.virtual_memory().available < 1.6 * nbytes:
            return {"value": None, "unit": "particle-steps/s", "h2d_bytes_per_step": nbytes,
                    "d2h_bytes_per_step": nbytes, "note": "not enough host memory for the pinned state"}
        names = "dx dy dz ux uy uz w".split()
        host = {m: torch.empty(n, dtype=torch.float32, pin_memory=True).numpy() for m in names}
        host["cell"] = torch.empty(n, dtype=torch.int32, pin_memory=True).numpy()
        hf = torch.empty((9, c.nc), dtype=torch.float32, pin_memory=True).numpy()
        import ctypes as C
        L = c.L
        ptrs = [host[m].ctypes.data_as(C.c_void_p) for m in names] + [host["cell"].ctypes.data_as(C.c_void_p)]
        fptr = (C.c_void_p * 9)(*[hf[m].ctypes.data for m in range(9)])
        got = C.c_int64()
        c._ck(L.cpic_download_particles(c.h, *ptrs, n, C.byref(got)))
        c._ck(L.cpic_download_fields(c.h, fptr))
        c.sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            # one call: particles stream host -> device -> host in chunks, both PCIe directions busy at once
            en = c.step_host(self.k, host, host, hf, hf, energies=True)
        sec = time.perf_counter() - t0
        return {"value": n * steps / sec, "unit": "particle-steps/s", "h2d_bytes_per_step": nbytes,
                "d2h_bytes_per_step": nbytes + 16, "steps": steps, "seconds": sec,
                "what": "per step: ONE cpic_step_host call on pinned host arrays (8 particle members + 9 field "
                        "components in, the same out, + energies): chunked H2D / in-place push / D2H pipeline, "
                        "then the field advance"}

    def close(self):
        self.c.close()


def side_lines(device):
    """Side lines for the other BASELINE configs (N = 1): C2 through the C ABI, C1 / C3 through the C++ host facade
    (examples/build/cbnpic_<deck>: the reference's own deck source, unmodified, on the facade headers) with the
    reference's CPU build of the same deck beside it.  Bounded to a few seconds each; failures are reported, not fatal."""
    import re
    out = {}
    # C2: the two-stream deck scaled to 1e8 particles on 32 cells, ES solver
    try:
        import cabanapic_b200 as cp
        from cabanapic_b200 import decks
        d = decks.two_stream_short(np.float32, orientation="x")
        d.nppc = 3125000
        k, _, we = d.consts()
        n = d.num_particles
        with cp.Context(d.nx, d.ny, d.nz, 1, max_particles=n, real=np.float32, solver=cp.SOLVER_ES_1D, device=device) as c:
            c.upload_particles(d.initial_particles())      # the deck's initialiser on the host, as the reference runs it
            c.upload_fields(d.initial_fields())
            c.step(k, 8, cp.SORT_FUSED, False); c.sync()
            c.step(k, 40, cp.SORT_FUSED, False); c.sync()
            ms = c.last_ms(3) / 40
        out["c2_two_stream_1e8_es"] = {"particles": n, "cells": [d.nx, d.ny, d.nz], "ms_per_step": ms,
                                       "value": n / (ms * 1e-3), "unit": "particle-steps/s", "path": "cpic_step(CPIC_SORT_FUSED)"}
    except Exception as ex:
        out["c2_two_stream_1e8_es"] = {"error": str(ex)[:300]}
    # C1 / C3 through the facade binaries
    from oracle.api import RefLib
    # (C4: the Weibel deck at 128^3 x 32 ppc = 6.7e7 particles through the facade's push<> / advance_* calls, default settings)
    for key, exe, deck, steps, extra_env in (("c1_two_stream_em_facade", "cbnpic_2stream-em", "2stream-em", 2000, {}),
                                             ("c3_dioctron_3d_facade", "cbnpic_dioctron_3d", "dioctron_3d", 2000, {}),
                                             ("c4_weibel_3d_facade", "cbnpic_weibel_3d", None, 24, {"CPIC_WEIBEL_N": "128", "CPIC_WEIBEL_PPC": "32"})):
        path = os.path.join(ROOT, "examples", "build", exe)
        try:
            if not os.path.exists(path):
                raise RuntimeError(f"{path} not built")
            env = dict(os.environ, CPIC_STEPS=str(steps), CPIC_ENERGY_INTERVAL="0", CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(device)))
            env.update(extra_env)
            r = subprocess.run([path], capture_output=True, text=True, timeout=300, env=env, cwd="/tmp")
            m = re.search(r"#(\d+) steps of (\d+) particles in ([0-9.]+) s: ([0-9.e+]+) particle-steps/s", r.stdout)
            if not m:
                raise RuntimeError("no summary line: " + (r.stdout[-200:] + r.stderr[-200:]))
            m2 = re.search(r"#steady: (\d+) steps after the first .* in ([0-9.]+) s: ([0-9.e+]+) particle-steps/s", r.stdout)
            e = {"steps": int(m.group(1)), "particles": int(m.group(2)), "seconds": float(m.group(3)),
                 "value": float(m.group(4)), "unit": "particle-steps/s", "steps_per_s": int(m.group(1)) / float(m.group(3)),
                 "path": f"examples/build/{exe} (C++ facade over the C ABI, default settings, wall clock incl. launches)"}
            if m2:      # without the first step, which carries the one-time upload of the host-initialised particles
                e["steady_value"] = float(m2.group(3))
                e["steady_steps_per_s"] = int(m2.group(1)) / max(float(m2.group(2)), 1e-9)
            if deck and RefLib.available(deck, "f32"):
                R = RefLib(deck, "f32").create_from_deck(solver=0)
                kk, _, _ = R.deck_consts()
                R.run(kk, 20)
                t0 = time.perf_counter(); R.run(kk, 200); el = time.perf_counter() - t0
                e["cpu_reference_steps_per_s"] = 200 / el
                e["speedup_vs_cpu_reference"] = e["steps_per_s"] / (200 / el)
            out[key] = e
        except Exception as ex:
            out[key] = {"error": str(ex)[:300]}
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

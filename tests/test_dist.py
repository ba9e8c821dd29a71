"""Multi-rank host logic on CPU (gloo, world_size 2): the slab and replicated steppers of
cabanapic_b200.dist driven by the oracle-backed engine, compared with the single-domain oracle.

What is pinned here: the exchange choreography (accumulator ghost rows, particle migration with
cell re-basing, the z sweeps of the J fold in the reference's order, J / cB ghost-copy planes),
the slab partition arithmetic and the migration counts.  The CUDA kernels behind the same calls
are pinned by tests/test_gpu_parity.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import PREC, canonical_order, consts_for, random_state
from oracle.api import PARTICLE_NAMES, Restatement


def test_slab_ranges_and_mode():
    from cabanapic_b200.dist import choose_mode, slab_ranges
    assert slab_ranges(256, 8) == [(32 * r, 32) for r in range(8)]
    r = slab_ranges(10, 4)
    assert [n for _, n in r] == [3, 3, 2, 2] and r[-1][0] + r[-1][1] == 10
    assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(3))
    assert choose_mode(256, 256, 256, 8) == "slab"
    assert choose_mode(1, 32, 1, 8) == "replicated"          # C1: tiny grid, replicate
    assert choose_mode(64, 64, 1, 4) == "replicated"         # C3: nz = 1 cannot be sliced
    assert choose_mode(32, 32, 32, 1) == "replicated"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _split_state(s, z0, nzl):
    """Local slab state: field planes [z0, z0+nzl+2) of the global arrays (so the ghost planes hold
    the neighbours' data, wrapping periodically through the global ghosts), the particles of the
    owned planes with re-based cell indices."""
    gx, gy = s.nx + 2, s.ny + 2
    plane = gx * gy
    f = s.f.reshape(9, s.nz + 2, plane)
    zsel = [(z0 + j - 1) % s.nz + 1 for j in range(nzl + 2)]      # global plane of local plane j
    zsel[0] = z0 if z0 > 0 else 0                                   # the global low ghost for rank 0
    zsel[-1] = z0 + nzl + 1                                         # == global high ghost for the last rank
    lf = f[:, zsel, :].reshape(9, -1).copy()
    iz = s.p["cell"] // plane
    sel = (iz >= z0 + 1) & (iz <= z0 + nzl)
    p = {k: s.p[k][sel].copy() for k in PARTICLE_NAMES}
    p["cell"] = (p["cell"] - z0 * plane).astype(np.int32)
    return lf, p, np.nonzero(sel)[0]


def _slab_worker(rank, world, port, prec, grid, nsteps, outdir, use_async=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cabanapic_b200._lib import Consts
        from cabanapic_b200.dist import SlabStepper, slab_ranges
        from slab_engine import OracleEngine
        nx, ny, nz = grid
        s = random_state(nx, ny, nz, nppc=12, prec=prec, seed=21)
        k = consts_for(nx, ny, nz, prec)
        ranges = slab_ranges(nz, world)
        z0, nzl = ranges[rank]
        lf, p, _ = _split_state(s, z0, nzl)
        e = OracleEngine(nx, ny, nzl, capacity=3 * len(p["cell"]) + 64, prec=prec, z_periodic=False, async_migration=use_async)
        e.s.f[:] = lf
        e.set_particles(p)
        st = SlabStepper(e, Consts(**k.to_dict()), rank, world, ranges[(rank - 1) % world][1],
                         ranges[(rank + 1) % world][1], send_capacity=len(p["cell"]) + 16)
        mig, en = [], []
        for _ in range(nsteps):
            st.step(fused=use_async)       # the device-counted exchange is what fused steps use when the engine has it
            mig.append(st.last_migration)
            en.append(st.energies())
        assert st._async_used == use_async if use_async else True
        out = e.particles()
        out["cell"] = out["cell"] + z0 * (nx + 2) * (ny + 2)       # back to global numbering
        np.savez(os.path.join(outdir, f"slab_{rank}.npz"), f=e.s.f, mig=np.array(mig), en=np.array(en),
                 z0=z0, nzl=nzl, **{"p_" + n: out[n] for n in PARTICLE_NAMES})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("use_async", [False, True])
@pytest.mark.parametrize("prec,tol", [("f64", 1e-12), ("f32", 2e-5)])
def test_slab_mode_matches_single_domain_oracle(tmp_path, prec, tol, use_async):
    """use_async: the one-batch exchange with the counts travelling as tensors (accumulator planes + counts +
    whole capacity-sized particle payloads in ONE group of sends/receives), as the GPU runs use it."""
    world, grid, nsteps = 2, (5, 4, 6), 6
    nx, ny, nz = grid
    mp.spawn(_slab_worker, args=(world, _free_port(), prec, grid, nsteps, str(tmp_path), use_async), nprocs=world, join=True)
    # single-domain oracle on the same state, tracking per-step slab crossings
    s = random_state(nx, ny, nz, nppc=12, prec=prec, seed=21)
    k = consts_for(nx, ny, nz, prec)
    O = Restatement(prec)
    plane = (nx + 2) * (ny + 2)
    from cabanapic_b200.dist import slab_ranges
    ranges = slab_ranges(nz, world)
    owner_of_plane = np.zeros(nz + 2, dtype=int)
    for r, (z0, n) in enumerate(ranges):
        owner_of_plane[z0 + 1: z0 + n + 1] = r
    want_mig, want_en = np.zeros((world, nsteps, 2), dtype=int), []
    for t in range(nsteps):
        iz0 = s.p["cell"] // plane
        want_en.append(O.step(s, k, 0, 1, energies=True)[0])
        iz1 = s.p["cell"] // plane
        o0, o1 = owner_of_plane[iz0], owner_of_plane[iz1]
        moved = o0 != o1
        # direction: +z (incl. the periodic wrap nz -> 1) is "hi", -z is "lo"
        up = moved & ((iz1 == iz0 + 1) | ((iz0 == nz) & (iz1 == 1)))
        dn = moved & ~up
        for r in range(world):
            want_mig[r, t] = [np.sum(dn & (o0 == r)), np.sum(up & (o0 == r))]
    got = [np.load(tmp_path / f"slab_{r}.npz") for r in range(world)]
    # migration counts: bit-exact
    for r in range(world):
        assert np.array_equal(got[r]["mig"], want_mig[r]), (r, got[r]["mig"], want_mig[r])
    assert want_mig.sum() > 20
    # particles: same multiset; cell indices bit-exact, state to the tolerance
    P = {n: np.concatenate([g["p_" + n] for g in got]) for n in PARTICLE_NAMES}
    assert len(P["cell"]) == s.np
    og, oo = canonical_order(P), canonical_order(s.p)
    assert np.array_equal(P["cell"][og], s.p["cell"][oo])
    for n in PARTICLE_NAMES[:7]:
        assert np.allclose(P[n][og], s.p[n][oo], rtol=0, atol=tol), n
    # fields: owned planes of every slab equal the global oracle's planes
    gf = s.f.reshape(9, nz + 2, plane)
    scale = np.abs(gf).max(axis=(1, 2), keepdims=True) + 1e-30
    for r, g in enumerate(got):
        z0, nzl = int(g["z0"]), int(g["nzl"])
        lf = g["f"].reshape(9, nzl + 2, plane)
        assert (np.abs(lf[:, 1:nzl + 1] - gf[:, z0 + 1:z0 + nzl + 1]) / scale).max() < tol
    # energies: sum over ranks equals the global history
    en = got[0]["en"]
    assert np.allclose(en, np.array(want_en), rtol=max(tol, 1e-6))


def _repl_worker(rank, world, port, prec, grid, nsteps, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cabanapic_b200._lib import Consts
        from cabanapic_b200.dist import ReplicatedStepper
        from slab_engine import OracleEngine
        nx, ny, nz = grid
        s = random_state(nx, ny, nz, nppc=10, prec=prec, seed=4)
        k = consts_for(nx, ny, nz, prec)
        lo, hi = s.np * rank // world, s.np * (rank + 1) // world
        e = OracleEngine(nx, ny, nz, capacity=hi - lo, prec=prec)
        e.s.f[:] = s.f
        e.set_particles({n: s.p[n][lo:hi] for n in PARTICLE_NAMES})
        st = ReplicatedStepper(e, Consts(**k.to_dict()), rank, world)
        for _ in range(nsteps):
            st.step()
        p = e.particles()
        np.savez(os.path.join(outdir, f"repl_{rank}.npz"), f=e.s.f, **{"p_" + n: p[n] for n in PARTICLE_NAMES})
    finally:
        dist.destroy_process_group()


def test_replicated_mode_matches_single_domain_oracle(tmp_path):
    world, grid, nsteps, prec = 2, (1, 16, 1), 10, "f64"
    nx, ny, nz = grid
    mp.spawn(_repl_worker, args=(world, _free_port(), prec, grid, nsteps, str(tmp_path)), nprocs=world, join=True)
    s = random_state(nx, ny, nz, nppc=10, prec=prec, seed=4)
    k = consts_for(nx, ny, nz, prec)
    Restatement(prec).step(s, k, 0, nsteps)
    got = [np.load(tmp_path / f"repl_{r}.npz") for r in range(world)]
    assert np.array_equal(got[0]["f"], got[1]["f"])              # replicas stay bit-identical
    scale = np.abs(s.f).max() + 1e-30
    assert np.abs(got[0]["f"] - s.f).max() / scale < 1e-12
    for n in PARTICLE_NAMES:                                     # particle order is preserved (no sort)
        P = np.concatenate([g["p_" + n] for g in got])
        if n == "cell":
            assert np.array_equal(P, s.p[n])
        else:
            assert np.allclose(P, s.p[n], rtol=0, atol=1e-12)

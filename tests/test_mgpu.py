"""GPU tests of the NATIVE multi-GPU layer (include/cabanapic_b200_mgpu.h: host C++ + NCCL inside the library) against
the single-domain oracle: per-(step, face) migration counts and cell indices bit-exact, particle state / fields /
energy history to summation order.  Runs on min(device_count, 2) ranks -- one process per GPU, the 128-byte NCCL id
handed over through cpic_mgpu_bootstrap_file; on a single-GPU box the same slab code runs on one rank whose z
neighbours are itself (CPIC_MGPU_OPEN_Z=1).  The box is a small C4-style plasma (3-D, anisotropic momenta, non-zero
fields) with more than 1024 cells per slab, so the float path is k_push3 (block-owned cell chunks)."""
import os

import numpy as np
import pytest

from helpers import canonical_order, consts_for, random_state
from oracle.api import PARTICLE_NAMES, Restatement

pytestmark = pytest.mark.gpu

NPPC, NSTEPS = 12, 7


def _grid():
    """8 planes per slab: more than 1024 cells each, so that every rank runs k_push3 (and the graph path applies)"""
    return (12, 10, 8 * max(2, _world()))


def _world():
    import torch
    return min(torch.cuda.device_count(), int(os.environ.get("CPIC_TEST_WORLD", "2")))


def _worker(rank, world, idfile, outdir, mode, graph):
    os.environ["CPIC_MGPU_OPEN_Z"] = "1"
    import ctypes as C
    import cabanapic_b200 as cp
    from cabanapic_b200 import _lib
    from test_dist import _split_state
    nx, ny, nz = _grid()
    s = random_state(nx, ny, nz, nppc=NPPC, prec="f32", seed=21)
    s.p["uz"] *= 0.5                                   # thermal anisotropy (C4: Weibel-type)
    k = cp.Consts(**consts_for(nx, ny, nz, "f32").to_dict())
    L = cp.lib()
    uid = C.create_string_buffer(128)
    assert L.cpic_mgpu_bootstrap_file(idfile.encode(), rank, world, 60.0, uid) == 0
    if mode == "slab":
        base, rem = divmod(nz, world)
        nzl = base + (1 if rank < rem else 0)
        z0 = rank * base + min(rank, rem)
        lf, p, _ = _split_state(s, z0, nzl)
        m = cp.Mgpu(nx, ny, nz, rank, world, uid.raw, mode=cp.MGPU_SLAB, max_particles=3 * len(p["cell"]) + 64, device=rank,
                    send_capacity=len(p["cell"]) + 16)
        assert (m.z0, m.nzl) == (z0, nzl)
    else:
        lo, hi = s.np * rank // world, s.np * (rank + 1) // world
        p, lf, z0, nzl = {n: s.p[n][lo:hi].copy() for n in PARTICLE_NAMES}, s.f, 0, nz
        m = cp.Mgpu(nx, ny, nz, rank, world, uid.raw, mode=cp.MGPU_REPLICATED, max_particles=hi - lo, device=rank)
    m.ctx.upload_particles(p)
    m.ctx.upload_fields(lf)
    transport = m.transport
    mig, en = [], []
    if graph == "host":                                # the slab lives in host memory: cpic_mgpu_step_host per step
        capn = 3 * len(p["cell"]) + 64
        a = {n: np.zeros(capn, dtype=p[n].dtype) for n in PARTICLE_NAMES}
        b = {n: np.zeros(capn, dtype=p[n].dtype) for n in PARTICLE_NAMES}
        cnt = len(p["cell"])
        for n in PARTICLE_NAMES:
            a[n][:cnt] = p[n]
        fa, fb = np.ascontiguousarray(lf, dtype=np.float32).copy(), np.zeros_like(lf, dtype=np.float32)
        used = False
        for _ in range(NSTEPS):
            cnt = m.step_host(k, a, b, cnt, fa, fb)
            mig.append(m.last_migration())
            a, b, fa, fb = b, a, fb, fa
        assert m.ctx.num_particles == cnt
        host_p = {n: a[n][:cnt].copy() for n in PARTICLE_NAMES}
        host_f = fa.copy()
    elif graph == "odd":                                 # an odd number of eager steps between two replays
        m.step(k, 2, cp.SORT_FUSED, use_graph=True)      # eager (the graph path needs two steps behind it)
        m.step(k, 2, cp.SORT_FUSED, use_graph=True)      # capture + replay
        used = m.used_graph
        m.step(k, 1, cp.SORT_FUSED, use_graph=True)      # eager: the other halves of the double buffers are current now
        m.step(k, NSTEPS - 5, cp.SORT_FUSED, use_graph=True)   # must capture again, not replay the stale graph
        used = used and m.used_graph
    elif graph:                                          # whole run in two calls: eager warm-up, then graph replay
        m.step(k, 3, cp.SORT_FUSED, use_graph=True)
        m.step(k, NSTEPS - 3, cp.SORT_FUSED, use_graph=True)
        used = m.used_graph
    else:
        used = False
        for _ in range(NSTEPS):
            m.step(k, 1, cp.SORT_FUSED)
            mig.append(m.last_migration())
            en.append(m.energies())
    dg = m.state_digest()
    tot = m.migration_counts()
    out = m.ctx.download_particles()
    f = m.ctx.download_fields()
    if graph == "host":                                # the host copy IS the result; it must also equal the device's state
        og, oo = canonical_order(host_p), canonical_order(out)
        for n in PARTICLE_NAMES:
            assert np.array_equal(host_p[n][og], out[n][oo]), n
        assert np.array_equal(host_f, f)
        out = host_p
    out["cell"] = out["cell"] + z0 * (nx + 2) * (ny + 2)
    m.close()
    np.savez(os.path.join(outdir, f"r{rank}.npz"), f=f, mig=np.array(mig), en=np.array(en), z0=z0, nzl=nzl, used=used, transport=transport,
             digest=np.array([dg[n] for n in _lib.DIGEST_NAMES]), tot=np.array(tot), **{"p_" + n: out[n] for n in PARTICLE_NAMES})


def _run(tmp_path, mode, graph):
    import torch.multiprocessing as mp
    world = _world()
    idfile = str(tmp_path / "nccl_id")
    mp.spawn(_worker, args=(world, idfile, str(tmp_path), mode, graph), nprocs=world, join=True)
    got = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    if mode == "slab" and world > 1:      # peer memory unless CPIC_MGPU_P2P=0 asks for NCCL (or IPC mapping is refused: stderr says so)
        tr = {str(g["transport"]) for g in got}
        print("slab transport:", tr)
        assert len(tr) == 1 and tr <= {"peer-memory", "nccl"}, tr
        if os.environ.get("CPIC_MGPU_P2P") == "0":
            assert tr == {"nccl"}
        if os.environ.get("CPIC_REQUIRE_P2P") == "1":
            assert tr == {"peer-memory"}
    return world, got


def _oracle(world):
    nx, ny, nz = _grid()
    s = random_state(nx, ny, nz, nppc=NPPC, prec="f32", seed=21)
    s.p["uz"] *= 0.5
    k = consts_for(nx, ny, nz, "f32")
    O = Restatement("f32")
    plane = (nx + 2) * (ny + 2)
    base, rem = divmod(nz, world)
    owner = np.zeros(nz + 2, dtype=int)
    for r in range(world):
        z0 = r * base + min(r, rem)
        owner[z0 + 1: z0 + base + (1 if r < rem else 0) + 1] = r
    mig, en = np.zeros((world, NSTEPS, 2), dtype=int), []
    for t in range(NSTEPS):
        iz0 = s.p["cell"] // plane
        en.append(O.step(s, k, 0, 1, energies=True)[0])
        iz1 = s.p["cell"] // plane
        # a particle migrates when it crosses a z face between slabs; with one rank every crossing of the periodic z
        # boundary (nz -> 1, 1 -> nz) is a migration to itself
        crossed_up = (iz1 == iz0 + 1) | ((iz0 == nz) & (iz1 == 1))
        crossed_dn = (iz1 == iz0 - 1) | ((iz0 == 1) & (iz1 == nz))
        left = (owner[iz0] != owner[iz1]) | ((iz0 == nz) & (iz1 == 1)) | ((iz0 == 1) & (iz1 == nz))
        for r in range(world):
            mig[r, t] = [np.sum(left & crossed_dn & (owner[iz0] == r)), np.sum(left & crossed_up & (owner[iz0] == r))]
    return s, mig, np.array(en)


def _check_state(world, got, s, tol=5e-5, ftol=2e-4):
    nx, ny, nz = _grid()
    plane = (nx + 2) * (ny + 2)
    P = {n: np.concatenate([g["p_" + n] for g in got]) for n in PARTICLE_NAMES}
    assert len(P["cell"]) == s.np
    og, oo = canonical_order(P), canonical_order(s.p)
    assert np.mean(P["cell"][og] == s.p["cell"][oo]) > 0.999          # float: a rounding-order difference may flip a crossing
    same = P["cell"][og] == s.p["cell"][oo]
    for n in PARTICLE_NAMES[:6]:
        assert np.abs(P[n][og][same] - s.p[n][oo][same]).max() < tol, n
    gf = s.f.reshape(9, nz + 2, plane)
    scale = np.abs(gf).max(axis=(1, 2), keepdims=True) + 1e-30
    for g in got:
        z0, nzl = int(g["z0"]), int(g["nzl"])
        lf = g["f"].reshape(9, nzl + 2, plane)
        assert (np.abs(lf[:, 1:nzl + 1] - gf[:, z0 + 1:z0 + nzl + 1]) / scale).max() < ftol


def test_native_slab_stepper_matches_oracle(tmp_path):
    """cpic_mgpu_step, eager, one step per call: migration counts per (step, face) bit-exact against the counts derived
    from the single-domain oracle, cells bit-exact, state / fields / energies to summation order."""
    world, got = _run(tmp_path, "slab", graph=False)
    s, want_mig, want_en = _oracle(world)
    for r in range(world):
        assert np.array_equal(got[r]["mig"], want_mig[r]), (r, got[r]["mig"], want_mig[r])
        assert np.array_equal(got[r]["tot"], want_mig[r].sum(axis=0))
    assert want_mig.sum() > 100
    _check_state(world, got, s)
    assert np.allclose(got[0]["en"], want_en, rtol=2e-4)              # all-reduced: every rank holds the global history
    d = got[0]["digest"]
    assert d[0] == s.np and d[2] == 0 and d[3] == 0 and d[7] == want_mig.sum()


def test_native_slab_stepper_graph_replay(tmp_path):
    """The same run as two calls with use_graph: pairs of steps replayed from a CUDA graph (NCCL kernels, device-counted
    migration and the external timing events included) end in the same state."""
    world, got = _run(tmp_path, "slab", graph=True)
    s, want_mig, _ = _oracle(world)
    assert bool(got[0]["used"]), "the graph path was not taken"
    for r in range(world):
        assert np.array_equal(got[r]["tot"], want_mig[r].sum(axis=0))
    _check_state(world, got, s)


def test_native_slab_stepper_graph_replay_after_odd_eager_steps(tmp_path):
    """A captured pair of steps bakes in which halves of the double buffers (particle store, segment bounds, cell counts)
    are current.  After an odd number of eager steps the graph must be captured again (CtxBase::state_signature), not
    replayed on stale buffers."""
    world, got = _run(tmp_path, "slab", graph="odd")
    s, want_mig, _ = _oracle(world)
    assert bool(got[0]["used"]), "the graph path was not taken"
    for r in range(world):
        assert np.array_equal(got[r]["tot"], want_mig[r].sum(axis=0))
    _check_state(world, got, s)


def test_native_slab_step_host_matches_oracle(tmp_path, monkeypatch):
    """cpic_mgpu_step_host: every rank keeps its slab in host arrays; per step the particles stream through the GPU in
    chunks (several per step here: CPIC_HOST_CHUNK), the exchanges and the field advance follow, and the host copy is
    patched where the migration changed the store.  Same bars as the device-resident stepper: migration counts per
    (step, face) bit-exact against the oracle, cells bit-exact, state / fields to summation order; the host copy equals
    the device store."""
    monkeypatch.setenv("CPIC_HOST_CHUNK", "1024")
    world, got = _run(tmp_path, "slab", graph="host")
    s, want_mig, _ = _oracle(world)
    for r in range(world):
        assert np.array_equal(got[r]["mig"], want_mig[r]), (r, got[r]["mig"], want_mig[r])
        assert np.array_equal(got[r]["tot"], want_mig[r].sum(axis=0))
    _check_state(world, got, s)


def test_native_replicated_stepper_matches_oracle(tmp_path):
    """REPLICATED mode: particles split by index, ncclAllReduce of the accumulator where the reference calls contribute."""
    world, got = _run(tmp_path, "replicated", graph=False)
    s, _, want_en = _oracle(world)
    for g in got[1:]:
        assert np.array_equal(g["f"], got[0]["f"])                    # replicas stay bit-identical
    _check_state(world, got, s)
    assert np.allclose(got[0]["en"], want_en, rtol=2e-4)


def test_weibel_deck_cpp_driver_on_gpus(tmp_path):
    """BASELINE configs[3] through the C++ host side: examples/cbnpic_mgpu.cpp (deck interface + cpic_mgpu_*, one
    process per GPU, file rendezvous) runs the Weibel deck in slab mode on min(device_count, 2) GPUs with particle
    migration; its energy history must agree with the single-GPU facade driver (cbnpic_weibel_3d) on the same deck,
    no particle may be lost, and particles must really have migrated."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    build = os.path.join(root, "examples", "build")
    if not os.path.exists(os.path.join(build, "cbnpic_mgpu_weibel_3d")):
        pytest.skip("examples/build was not produced in the build container")
    world = _world()
    env = dict(os.environ, CPIC_WEIBEL_N="24", CPIC_WEIBEL_PPC="8", CPIC_STEPS="40", CPIC_ENERGY_INTERVAL="10")
    single = tmp_path / "single"
    single.mkdir()
    r = subprocess.run([os.path.join(build, "cbnpic_weibel_3d")], cwd=single, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    want = np.loadtxt(single / "energies.txt", ndmin=2)
    multi = tmp_path / "multi"
    multi.mkdir()
    procs = []
    for rank in range(world):
        e = dict(env, CPIC_WORLD=str(world), CPIC_RANK=str(rank), CPIC_MGPU_ID_FILE=str(multi / "id"), CPIC_MGPU_MODE="slab",
                 CPIC_MGPU_OPEN_Z="1")
        procs.append(subprocess.Popen([os.path.join(build, "cbnpic_mgpu_weibel_3d")], cwd=multi, env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, err) in zip(procs, outs):
        assert p.returncode == 0, err[-2000:]
    got = np.loadtxt(multi / "energies.txt", ndmin=2)
    assert got.shape == want.shape == (4, 4)
    assert np.allclose(got[:, 2:], want[:, 2:], rtol=2e-3), (got, want)
    line = [l for l in outs[0][0].splitlines() if l.startswith("#digest:")][0].split()
    d = dict(zip(line[1::2], line[2::2]))
    n = 24 ** 3 * 8
    assert float(d["particles"]) == n and float(d["not-interior"]) == 0 and float(d["offsets-out"]) == 0
    assert float(d["migrated"]) > 1000
    assert "z-slab" in outs[0][0]

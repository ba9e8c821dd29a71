"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle.

Bars (stated per test):
  * index/integer work (cell indices, mover/crossing/wrap counts, sort permutation): bit-exact;
  * field-side kernels (interpolator load, unload, ghost fold/copy, advance_b/e): bit-exact --
    they evaluate the reference's expressions in the reference's order;
  * per-particle state after push in CPIC_FP_STRICT: bit-exact (same inputs -> same bits);
  * accumulators: same sums in a different association (warp tree / atomics) ->
    float 2e-5 relative to the row scale, double 1e-12;
  * CPIC_FP_CONTRACT (fused multiply-add): float 2e-6 / double 1e-14 absolute on O(1) state;
  * multi-step histories: the tolerance written in each test.
"""
import os

import numpy as np
import pytest

from helpers import PREC, canonical_order, consts_for, random_state
from oracle.api import PARTICLE_NAMES, Consts as OConsts, Restatement, State

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def cp():
    import cabanapic_b200 as m
    return m


def to_k(k):
    """oracle Consts -> product Consts (same field layout, different ctypes class)."""
    return cp().Consts(**k.to_dict())


def make_ctx(s, **kw):
    m = cp()
    c = m.Context(s.nx, s.ny, s.nz, s.ng, max_particles=s.np + 100, real=PREC[s.prec], **kw)
    c.upload_particles(s.p)
    c.upload_fields(s.f)
    return c


def acc_close(a, b, prec):
    scale = np.abs(b).max() + 1e-30
    tol = 2e-5 if prec == "f32" else 1e-12
    return np.abs(a - b).max() <= tol * scale


GRIDS = [(6, 5, 4), (1, 7, 3), (4, 1, 1), (1, 32, 1), (3, 3, 1), (17, 9, 5)]


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("grid", GRIDS)
def test_field_kernels_bitwise(grid, prec):
    nx, ny, nz = grid
    s = random_state(nx, ny, nz, nppc=4, prec=prec, seed=5)
    rng = np.random.default_rng(1)
    s.acc[:] = rng.standard_normal(s.acc.shape).astype(PREC[prec])
    s.f[6:] = rng.standard_normal(s.f[6:].shape).astype(PREC[prec])   # junk J incl. ghosts
    k = consts_for(nx, ny, nz, prec)
    O = Restatement(prec)
    with make_ctx(s) as c:
        c.upload_accumulators(s.acc)
        O.load_interpolator(s); c.load_interpolator_array()
        assert np.array_equal(c.download_interpolators(), s.interp)
        # ghost fold / copy on junk J
        O.ghost_fold(s, (6, 7, 8)); c.update_ghosts(0)
        assert np.array_equal(c.download_fields(), s.f)
        O.ghost_copy(s, (6, 7, 8)); c.update_ghosts(1)
        assert np.array_equal(c.download_fields(), s.f)
        O.unload_accumulator(s, k); c.unload_accumulator_array(to_k(k))
        assert np.array_equal(c.download_fields(), s.f)
        hp = (0.5 * k.px, 0.5 * k.py, 0.5 * k.pz)
        for _ in range(2):
            O.advance_b(s, *hp); c.advance_b(*hp)
            assert np.array_equal(c.download_fields(), s.f)
            O.advance_e(s, k.px, k.py, k.pz, k.dt_eps0); c.advance_e(k.px, k.py, k.pz, k.dt_eps0)
            assert np.array_equal(c.download_fields(), s.f)
        e, b = c.energies()
        eo, bo = O.energies(s)
        assert abs(e - eo) <= 1e-5 * eo and abs(b - bo) <= 1e-5 * bo


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_es1d_field_advance_bitwise(prec):
    s = random_state(24, 1, 1, nppc=4, prec=prec, seed=9)
    rng = np.random.default_rng(2)
    s.f[6:] = rng.standard_normal(s.f[6:].shape).astype(PREC[prec])
    k = consts_for(24, 1, 1, prec)
    O = Restatement(prec)
    with make_ctx(s, solver=cp().SOLVER_ES_1D) as c:
        for _ in range(3):
            O.advance_e(s, k.px, k.py, k.pz, k.dt_eps0, solver=1)
            c.advance_b(0.1, 0.1, 0.1)   # ES: no-op
            c.advance_e(k.px, k.py, k.pz, k.dt_eps0)
            assert np.array_equal(c.download_fields(), s.f)
        e, b = c.energies()
        eo, _ = O.energies(s, solver=1)
        assert abs(e - eo) <= 1e-5 * eo and b == 0.0


@pytest.mark.parametrize("deposit", [1, 2, 3])
@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("grid", GRIDS)
def test_push_strict_bitwise_teacher_forced(grid, prec, deposit):
    """One push from identical inputs: particle members and cell indices bit-exact, mover /
    crossing counts exact, accumulators equal up to summation order."""
    nx, ny, nz = grid
    s = random_state(nx, ny, nz, nppc=37, prec=prec, seed=21)
    k = consts_for(nx, ny, nz, prec)
    O = Restatement(prec)
    with make_ctx(s, deposit_mode=deposit) as c:
        c.enable_push_stats(True)
        O.load_interpolator(s); c.load_interpolator_array()
        O.clear_accumulator(s); c.clear_accumulator_array()
        movers, crossings = O.push(s, k)
        c.push(to_k(k))
        p = c.download_particles()
        for n in PARTICLE_NAMES:
            assert np.array_equal(p[n], s.p[n]), n
        st = c.push_stats()
        assert st["movers"] == movers and st["crossings"] == crossings
        assert movers > 0
        assert acc_close(c.download_accumulators(), s.acc, prec)


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_push_sorted_warp_uniform_path(prec):
    """Cell-sorted particles with many per cell exercise the one-row-per-warp fast path."""
    s = random_state(5, 4, 3, nppc=200, prec=prec, seed=33, uth=0.05)
    k = consts_for(5, 4, 3, prec)
    O = Restatement(prec)
    with make_ctx(s, deposit_mode=3) as c:
        O.load_interpolator(s); c.load_interpolator_array()
        O.clear_accumulator(s); c.clear_accumulator_array()
        O.push(s, k); c.push(to_k(k))
        p = c.download_particles()
        for n in PARTICLE_NAMES:
            assert np.array_equal(p[n], s.p[n]), n
        assert acc_close(c.download_accumulators(), s.acc, prec)


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_push_contract_mode_close(prec):
    s = random_state(6, 5, 4, nppc=50, prec=prec, seed=4)
    k = consts_for(6, 5, 4, prec)
    O = Restatement(prec)
    with make_ctx(s, fp_mode=cp().FP_CONTRACT) as c:
        O.load_interpolator(s); c.load_interpolator_array()
        O.clear_accumulator(s); c.clear_accumulator_array()
        O.push(s, k); c.push(to_k(k))
        p = c.download_particles()
        tol = 2e-6 if prec == "f32" else 1e-14
        same_cell = p["cell"] == s.p["cell"]
        assert same_cell.mean() > 0.999            # a face tie can flip under different rounding
        for n in ("dx", "dy", "dz", "ux", "uy", "uz"):
            assert np.abs(p[n][same_cell] - s.p[n][same_cell]).max() < tol, n
        a, b = c.download_accumulators(), s.acc
        assert np.abs(a - b).max() <= (1e-4 if prec == "f32" else 1e-11) * np.abs(b).max()


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_uncenter_bitwise(prec):
    s = random_state(6, 5, 4, nppc=20, prec=prec, seed=8)
    k = consts_for(6, 5, 4, prec)
    O = Restatement(prec)
    with make_ctx(s) as c:
        O.load_interpolator(s); c.load_interpolator_array()
        O.uncenter(s, k.qdt_2mc); c.uncenter_particles(k.qdt_2mc)
        p = c.download_particles()
        for n in PARTICLE_NAMES:
            assert np.array_equal(p[n], s.p[n]), n


@pytest.mark.parametrize("name,ptol,etol", [
    ("2stream-em_f32", 2e-4, 2e-3), ("2stream-em_f64", 1e-11, 1e-10), ("custom_init_f32", 2e-4, 1e-2),
    ("dioctron_3d_f32", 2e-4, 2e-4), ("2particle_f32", 1e-4, 1e-3), ("random3d_f32", 5e-4, 1e-4),
    ("random3d_f64", 1e-11, 1e-11)])
def test_golden_fixtures_multi_step(name, ptol, etol):
    """Run the reference-generated fixtures for their N steps through the fused step call.
    Deposit association differs from the serial reference, so fields (and through them the
    particles) drift at rounding level: cells must agree for > 99.5 % of particles, positions /
    momenta within ptol absolute where the cell agrees, energies within etol relative.
    (custom_init in float: the whole 30-step history sits at the rounding-noise floor of the cancelling beam
    currents, E energy 7e-11; the atomic order of a run alone moves the figure between 0.8e-3 and 3e-3 -- measured
    over 16 runs of both deposit targets, global reductions and the block-private accumulator -- hence 1e-2.)"""
    z = np.load(os.path.join(GOLDEN, f"state_{name}.npz"))
    prec = name.split("_")[-1]
    nx, ny, nz, ng, nsteps, solver = [int(v) for v in z["meta"]]
    s = State(nx, ny, nz, ng, len(z["p0_cell"]), prec)
    for n in PARTICLE_NAMES:
        s.p[n][:] = z["p0_" + n]
    s.f[:] = z["f0"]
    k = cp().Consts(**dict(zip("qdt_2mc cdt_dx cdt_dy cdt_dz qsp dx dy dz dt px py pz dt_eps0".split(),
                               [float(v) for v in z["consts"]])))
    with make_ctx(s, solver=solver) as c:
        en = c.step(k, nsteps, sort_interval=0, energies=True)
        p = c.download_particles()
        f = c.download_fields()
    same = p["cell"] == z["p1_cell"]
    assert same.mean() > 0.995
    for n in ("dx", "dy", "dz", "ux", "uy", "uz"):
        assert np.abs(p[n][same] - z["p1_" + n][same]).max() < ptol, n
    ge = z["energies"]
    # energies relative to the history's own scale (early lines sit at the noise floor)
    assert (np.abs(en - ge).max(axis=0) / (ge.max(axis=0) + 1e-300)).max() < etol
    # E and cB against the field group's own scale.  (J is re-assigned from the accumulators every
    # step and is a cancellation residue of counter-streaming currents; its single-step parity is
    # covered by the accumulator checks above.)
    for grp in (slice(0, 3), slice(3, 6)):
        scale = np.abs(z["f1"][grp]).max() + 1e-30
        # absolute floor: in the 1-D two-stream decks E starts at the rounding noise of the
        # cancelling beam currents (|E| ~ 3e-7 in float), where summation order is everything
        floor = 1e-7 if prec == "f32" else 1e-15
        err = np.abs(f[grp] - z["f1"][grp]).max()
        assert err < 50 * ptol * scale + floor, (grp, err, scale)


def test_fused_step_equals_unfused_calls():
    """cpic_step is exactly the reference-named calls in the reference's order."""
    m = cp()
    from cabanapic_b200 import decks
    d = decks.custom_init(np.float32)
    a = m.Simulation(d, deposit_mode=1)
    b = m.Simulation(d, deposit_mode=1)
    for _ in range(5):
        a.step_unfused()
    b.run(5, energies=False)
    pa, pb = a.particles(), b.particles()
    # scalar atomics in arbitrary order: not bitwise, but tight
    for n in ("dx", "ux"):
        assert np.abs(pa[n] - pb[n]).max() < 1e-5
    assert np.array_equal(pa["cell"], pb["cell"])
    a.close(); b.close()


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("chunk", [0, 192])
def test_step_host_streamed_equals_oracle_step(prec, chunk, monkeypatch):
    """cpic_step_host (host-resident state, particles streamed through the device in chunks) does what the
    reference's loop body does (example/example.cpp:221-266): particles bit-exact after each of two steps
    (every chunk boundary incl. a ragged last chunk when chunk=192), fields and energies to accumulation-order
    tolerance; `out` aliasing `in`, and the context left holding the advanced state."""
    if chunk:
        monkeypatch.setenv("CPIC_HOST_CHUNK", str(chunk))
    R = PREC[prec]
    s = random_state(7, 5, 4, nppc=13, prec=prec, seed=21)           # 1820 particles: 9 full chunks + 92
    k = consts_for(7, 5, 4, prec)
    O = Restatement(prec)
    m = cp()
    with m.Context(s.nx, s.ny, s.nz, 1, max_particles=s.np, real=R) as c:
        p = {n: s.p[n].copy() for n in PARTICLE_NAMES}
        f = s.f.copy()
        for it in range(2):
            en_ref = O.step(s, k, 0, 1, energies=True)[0]
            f_out = np.empty_like(f)
            en = c.step_host(to_k(k), p, p, f, f_out, energies=True)   # in place on the host arrays
            for n in PARTICLE_NAMES:
                assert np.array_equal(p[n], s.p[n]), (it, n)
            scale = np.abs(s.f).max()
            assert np.abs(f_out - s.f).max() <= (2e-5 if prec == "f32" else 1e-12) * scale
            assert np.allclose(en, en_ref, rtol=1e-4 if prec == "f32" else 1e-10)
            # feed the oracle's fields back so that the second step starts from bit-identical inputs
            f = s.f.copy()
        q = c.download_particles()
        for n in PARTICLE_NAMES:
            assert np.array_equal(q[n], s.p[n]), n
        # no download requested: state advances on the device only
        O.step(s, k, 0, 1)
        c.step_host(to_k(k), p, None, f, None)
        q = c.download_particles()
        for n in PARTICLE_NAMES:
            assert np.array_equal(q[n], s.p[n]), n
        # a bad cell index is reported, not dereferenced
        p["cell"][7] = 10 ** 6
        with pytest.raises(m.CpicError) as e:
            c.step_host(to_k(k), p, p, f, None)
        assert e.value.code == -5


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_sort_particles_properties(prec):
    """Sort = a permutation, cells non-decreasing, idempotent, and physics-neutral."""
    s = random_state(7, 6, 5, nppc=30, prec=prec, seed=12)
    rng = np.random.default_rng(0)
    perm = rng.permutation(s.np)
    for n in PARTICLE_NAMES:
        s.p[n][:] = s.p[n][perm]
    k = consts_for(7, 6, 5, prec)
    with make_ctx(s) as c:
        c.sort_particles()
        p1 = c.download_particles()
        assert np.all(np.diff(p1["cell"]) >= 0)
        o0, o1 = canonical_order(s.p), canonical_order(p1)
        for n in PARTICLE_NAMES:
            assert np.array_equal(s.p[n][o0], p1[n][o1]), n
        c.sort_particles()
        p2 = c.download_particles()
        assert np.array_equal(p2["cell"], p1["cell"])
        # push after sort == push before sort, particle by particle
        O = Restatement(prec)
        O.load_interpolator(s); O.clear_accumulator(s); O.push(s, k)
        c.load_interpolator_array(); c.clear_accumulator_array(); c.push(to_k(k))
        p3 = c.download_particles()
        o0, o3 = canonical_order(s.p), canonical_order(p3)
        for n in PARTICLE_NAMES:
            assert np.array_equal(s.p[n][o0], p3[n][o3]), n
        assert acc_close(c.download_accumulators(), s.acc, prec)


def test_empty_and_ragged_inputs():
    m = cp()
    with m.Context(4, 3, 2, 1, max_particles=0) as c:     # no particles at all
        k = to_k(consts_for(4, 3, 2))
        c.upload_fields(np.zeros((9, c.nc), np.float32))
        c.step(k, 3, sort_interval=1, energies=True)
        assert c.num_particles == 0
    s = random_state(3, 2, 2, nppc=1, seed=2)             # 12 particles: one ragged warp
    for cut in (1, 5, 12):
        sub = {n: s.p[n][:cut].copy() for n in PARTICLE_NAMES}
        t = State(3, 2, 2, 1, cut, "f32")
        for n in PARTICLE_NAMES:
            t.p[n][:] = sub[n]
        t.f[:] = s.f
        O = Restatement("f32")
        kk = consts_for(3, 2, 2)
        with make_ctx(t) as c:
            O.load_interpolator(t); O.clear_accumulator(t); O.push(t, kk)
            c.load_interpolator_array(); c.clear_accumulator_array(); c.push(to_k(kk))
            p = c.download_particles()
            for n in PARTICLE_NAMES:
                assert np.array_equal(p[n], t.p[n])
            assert acc_close(c.download_accumulators(), t.acc, "f32")


def test_error_behaviour():
    m = cp()
    with pytest.raises(m.CpicError) as e:
        m.Context(4, 1, 1, 1, boundary=m.BOUNDARY_REFLECT, solver=m.SOLVER_ES_1D)
    assert e.value.code == -6                                 # Reflect is built for the EM solver only
    with pytest.raises(m.CpicError):
        m.Context(4, 4, 4, 2)                                 # ng != 1
    with pytest.raises(m.CpicError):
        m.Context(4, 4, 1, 1, solver=m.SOLVER_ES_1D)          # ES is 1-D only
    s = random_state(3, 3, 3, nppc=2)
    s.p["cell"][5] = 10 ** 6                                  # the decks/2stream-short.cxx failure mode
    with m.Context(3, 3, 3, 1, max_particles=100) as c:
        with pytest.raises(m.CpicError) as e:
            c.upload_particles(s.p)
        assert e.value.code == -5
    with m.Context(3, 3, 3, 1, max_particles=10) as c:
        with pytest.raises(m.CpicError) as e:
            c.upload_particles(random_state(3, 3, 3, nppc=20).p)
        assert e.value.code == -4


def test_current_conservation_at_scale():
    """Size-independent property at a BASELINE-like size (64^3 x 16 = 4.2 M particles, float):
    for every streak the four quadrant currents of a component sum to 4*q*(half displacement),
    so sum_cells sum_k acc[c][X][k] == 4 * sum_p q * u_X * cdt_dX / gamma, crossings or not."""
    from cabanapic_b200 import decks
    m = cp()
    d = decks.uniform_plasma(64, 64, 64, 16)
    k, _, we = d.consts()
    p0 = d.initial_particles()
    with m.Context(64, 64, 64, 1, max_particles=d.num_particles, real=np.float32) as c:
        c.upload_particles(p0)
        c.upload_fields(d.initial_fields())
        c.load_interpolator_array(); c.clear_accumulator_array()
        c.enable_push_stats(True)
        c.push(k)
        acc = c.download_accumulators().astype(np.float64).reshape(-1, 3, 4)
        p1 = c.download_particles()
        st = c.push_stats()
    u = np.stack([p1["ux"], p1["uy"], p1["uz"]]).astype(np.float64)
    gam = np.sqrt(1.0 + (u * u).sum(axis=0))
    q = p1["w"].astype(np.float64) * k.qsp
    for X, cdt in enumerate((k.cdt_dx, k.cdt_dy, k.cdt_dz)):
        want = 4.0 * np.sum(q * u[X] * cdt / gam)
        got = acc[:, X, :].sum()
        scale = 4.0 * np.sum(np.abs(q * u[X] * cdt / gam))
        assert abs(got - want) < 2e-6 * scale
    # E = B = 0: momenta unchanged bit for bit, ~10-16 % of particles change cell
    for n in ("ux", "uy", "uz", "w"):
        assert np.array_equal(p0[n], p1[n])
    frac = st["movers"] / d.num_particles
    assert 0.05 < frac < 0.25
    ix = p1["cell"] % 66; iy = (p1["cell"] // 66) % 66; iz = p1["cell"] // (66 * 66)
    assert ix.min() >= 1 and ix.max() <= 64 and iy.min() >= 1 and iy.max() <= 64 and iz.min() >= 1 and iz.max() <= 64
    assert np.abs(p1["dx"]).max() <= 1 and np.abs(p1["dy"]).max() <= 1 and np.abs(p1["dz"]).max() <= 1


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_kinetic_energy_diagnostic(prec):
    """cpic_kinetic_energy (sum w (gamma - 1), double accumulation; the reference has no such diagnostic,
    SURVEY 8f.1) against numpy in float64 on a random state incl. cold and relativistic particles."""
    s = random_state(6, 5, 4, nppc=50, prec=prec, seed=17)
    s.p["ux"][:10] = 0; s.p["uy"][:10] = 0; s.p["uz"][:10] = 1e-6      # cold: gamma - 1 ~ 5e-13
    s.p["ux"][10:20] = 30.0                                              # relativistic
    with make_ctx(s) as c:
        got = c.kinetic_energy()
    u2 = sum(s.p[n].astype(np.float64) ** 2 for n in ("ux", "uy", "uz"))
    want = float(np.sum(s.p["w"].astype(np.float64) * (u2 / (np.sqrt(1.0 + u2) + 1.0))))
    assert abs(got - want) <= 1e-12 * want


def test_energy_budget_is_conserved_at_scale():
    """Size-independent property of the whole loop (push + deposit + field advance) on the bench plasma at 64^3
    cells x 16 ppc = 4.2 M particles, 60 reordering steps: kinetic energy sum w (gamma-1) m c^2 plus field energy
    (eps0/2) sum (E^2 + cB^2) dV stays constant while the fields grow out of the deposit noise at the particles'
    expense.  Observed drift 4e-5 of the kinetic energy (explicit-PIC grid heating at dx = lambda_D); bar 5e-4."""
    m = cp()
    from cabanapic_b200 import decks
    d = decks.uniform_plasma(64, 64, 64, 16)
    k, _, we = d.consts()
    n = d.num_particles
    dV = k.dx * k.dy * k.dz
    with m.Context(64, 64, 64, 1, max_particles=n, real=np.float32) as c:
        c.init_uniform_plasma(0, n, 64, 64, 64, 16, weight=we)
        c.upload_fields(d.initial_fields())
        ke0 = c.kinetic_energy()
        e0, b0 = c.energies()
        assert e0 == 0.0 and b0 == 0.0
        tot = []
        for _ in range(6):
            c.step(k, 10, m.SORT_FUSED, False)
            e, b = c.energies()
            tot.append((c.kinetic_energy(), (e + b) * dV))
        ke1, fe1 = tot[-1]
        print(f"KE0 {ke0:.6e}  KE {ke1:.6e}  field {fe1:.6e}  drift {(ke1 + fe1 - ke0) / ke0:.3e}")
        assert fe1 > 0 and fe1 < 0.05 * ke0
        assert ke1 < ke0                                           # the particles paid for the fields
        for ke, fe in tot:
            assert abs(ke + fe - ke0) <= 5e-4 * ke0


def test_full_size_c5_properties():
    """BASELINE configs[4] at FULL size (256^3 cells x 64 ppc = 2^30 particles, 69 GB of particle store) through
    size-independent properties, all evaluated on the device.  The weights are first made a tag (2^20 classes of
    1024 particles, w = we*(1 + class/2^21); the push only ever multiplies w by the species charge).  After 4 steps
    of the default (reordering) step: the particle count is unchanged, every cell index is an interior voxel, every
    offset lies in [-1, 1], and the class histogram of the weights is unchanged -- no particle lost, duplicated or
    torn by the out-of-place slot claims, which is what a wrong cell histogram or an overlapping segment would do.
    Skipped down to the largest z extent the device's free memory allows."""
    import torch
    m = cp()
    from cabanapic_b200 import decks
    free, _ = torch.cuda.mem_get_info()
    nz = 256
    while nz > 16 and 256 * 256 * nz * 64 * 64 * 1.25 + 6e9 > free:
        nz //= 2
    if nz < 256:
        print(f"(device memory allows only nz = {nz})")
    d = decks.uniform_plasma(256, 256, nz, 64)
    k, _, we = d.consts()
    n = d.num_particles
    NCLS = 1 << 20
    chunk = 1 << 26

    class _Dev:
        def __init__(self, p, nwords):
            self.__cuda_array_interface__ = {"shape": (nwords,), "typestr": "<f4", "data": (int(p), False), "version": 2, "strides": None}

    with m.Context(256, 256, nz, 1, max_particles=n, real=np.float32) as c:
        c.init_uniform_plasma(0, n, 256, 256, nz, 64, weight=we)
        c.upload_fields(d.initial_fields())
        rec = lambda: torch.as_tensor(_Dev(c.device_ptr(0)[0], n * 8), device="cuda").view(n, 8)   # dx dy dz cell ux uy uz w
        c.sync()
        r = rec()
        for first in range(0, n, chunk):
            idx = torch.arange(first, min(first + chunk, n), device="cuda", dtype=torch.int64) % NCLS
            r[first:first + chunk, 7] = (np.float32(we) * (1.0 + idx.to(torch.float64) / (2.0 * NCLS))).to(torch.float32)
        torch.cuda.synchronize()

        def classes():
            r = rec()
            h = torch.zeros(NCLS, dtype=torch.int64, device="cuda")
            for first in range(0, n, chunk):
                w = r[first:first + chunk, 7].to(torch.float64)
                kcls = torch.round((w / float(np.float32(we)) - 1.0) * (2.0 * NCLS)).to(torch.int64)
                assert int(kcls.min()) >= 0 and int(kcls.max()) < NCLS
                h += torch.bincount(kcls, minlength=NCLS)
            return h
        h0 = classes()
        assert int(h0.sum()) == n
        c.step(k, 4, m.SORT_FUSED, False)
        c.sync()
        assert c.num_particles == n
        r = rec()
        gx, gy = 258, 258
        for first in range(0, n, chunk):
            rr = r[first:first + chunk]
            cell = rr[:, 3].contiguous().view(torch.int32)
            ix, iy, iz = cell % gx, (cell // gx) % gy, cell // (gx * gy)
            assert int(ix.min()) >= 1 and int(ix.max()) <= 256 and int(iy.min()) >= 1 and int(iy.max()) <= 256
            assert int(iz.min()) >= 1 and int(iz.max()) <= nz
            assert float(rr[:, :3].abs().max()) <= 1.0
            del cell, ix, iy, iz
        assert torch.equal(classes(), h0)


def test_full_size_c2_properties():
    """BASELINE configs[1] at full size: the two-stream deck (decks/2stream-short.cxx physics, x-oriented
    initialiser of decks/custom_init.cxx) scaled to 1e8 particles on 32 cells, ES field solver
    (-DSOLVER_TYPE=ES), 24 steps of the default step (block-private accumulator; CPIC_SORT_FUSED resolves to the
    periodic sort on such a grid).  Properties: particle count unchanged, every particle in an interior x cell with
    its offset in [-1, 1], energies finite, no B energy (ES), and the field energy grows out of the deposit noise
    (the two-stream instability) instead of staying at or blowing past it."""
    m = cp()
    from cabanapic_b200 import decks
    d = decks.two_stream_short(np.float32, "x")
    d.nppc = 3_125_000
    sim = m.Simulation(d, solver=m.SOLVER_ES_1D)
    try:
        n = sim.ctx.num_particles
        assert n == 100_000_000
        en = sim.run(24, sort_interval=m.SORT_FUSED, energies=True)
        assert np.all(np.isfinite(en)) and np.all(en[:, 1] == 0.0)
        assert en[-1, 0] > 10 * en[0, 0] and en[-1, 0] < 1e-3
        assert sim.ctx.num_particles == n
        p = sim.particles()
        ix = p["cell"] % (d.nx + 2)
        assert ix.min() >= 1 and ix.max() <= d.nx
        assert np.array_equal(p["cell"] // (d.nx + 2), np.full(n, 4, dtype=p["cell"].dtype))      # y = z = 1: (1 + 3*1)
        for a_ in ("dx", "dy", "dz"):
            assert np.abs(p[a_]).max() <= 1.0
    finally:
        sim.close()


def test_energy_history_2stream_em_double_vs_gold():
    """The reference's regression test (tests/energy_comparison): 6000 steps of the 1x32x1 EM
    two-stream deck in double.  Reference criterion: < 10 % on lines 3581..4880.  Ours: the
    double history must match the reference's gold file to 1e-4 on every sampled line."""
    from cabanapic_b200 import decks
    m = cp()
    gold = np.load(os.path.join(GOLDEN, "energies_gold_2stream-em.npz"))
    lines, g = gold["lines"], gold["f64"]
    sim = m.Simulation(decks.two_stream_em(np.float64))
    en = sim.run(6000, energies=True)[lines]
    sim.close()
    rel = np.abs(en - g) / np.minimum(en, g)
    window = (lines >= 3581) & (lines < 4881)
    assert rel[window].max() < 0.10
    assert rel.max() < 1e-4


def test_energy_history_2stream_em_float_vs_gold():
    """Float: chaotic after saturation.  The deposit sums in a different order than
    the serial reference, which perturbs the float history at rounding level.  In the linear
    phase the E energy oscillates between ~2e-11 and ~2e-16: at the minima the value IS the
    summation-order noise (merely permuting the particles in the serial oracle moves those
    lines by 2 %, tests/test_oracle.py::test_float_history_summation_order_sensitivity), so
    before saturation differences are judged against the local oscillation envelope (running
    max over +-64 lines of the gold file): < 0.5 % of the envelope (measured 0.02-0.09 % for
    every deposit mode).  Inside the reference's comparison window (lines 3581..4880,
    tests/energy_comparison/2stream-em.cxx:23,45-70) the reference asks |A-B|/min(A,B) < 10 %;
    its own serial code reaches 13.7 % there after a mere particle permutation (same oracle
    test), and the GPU's atomic order varies run to run (measured 4.5-8.4 %), so float is held
    to 25 % in the window.  Double (what the reference's CI runs) keeps 1e-4 on every line."""
    from cabanapic_b200 import decks
    m = cp()
    gold = np.load(os.path.join(GOLDEN, "energies_gold_2stream-em.npz"))
    lines, g, env = gold["lines"], gold["f32"], gold["env_f32"]
    sim = m.Simulation(decks.two_stream_em(np.float32))
    en = sim.run(6000, energies=True)[lines]
    sim.close()
    rel = np.abs(en - g) / np.minimum(en, g)
    erel = np.abs(en - g) / env
    window = (lines >= 3581) & (lines < 4881)
    assert erel[lines < 3581].max() < 5e-3, erel[lines < 3581].max()
    assert rel[window].max() < 0.25, rel[window].max()


# ----------------------------------------------------------------------------- slab-mode pieces
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_slab_mode_kernels_vs_oracle(prec):
    """The single-GPU pieces of the z-slab mode (cabanapic_b200/dist.py): push with the z wrap
    switched off (particles stay in the ghost planes; x/y wraps still apply there), the x/y-only
    ghost fold sweeps and ghost copy, the bare stencils, and the extraction of the ghost-plane
    particles with cell re-basing.  Bars: particles and fields bit-exact, accumulators to
    summation order, the extracted sets equal as multisets."""
    import torch
    nx, ny, nz = 6, 5, 4
    s = random_state(nx, ny, nz, nppc=16, prec=prec, seed=9)
    rng = np.random.default_rng(2)
    s.f[6:] = rng.standard_normal(s.f[6:].shape).astype(PREC[prec])
    k = consts_for(nx, ny, nz, prec)
    O = Restatement(prec)
    plane = (nx + 2) * (ny + 2)
    with make_ctx(s) as c:
        c.set_axis_periodic(1, 1, 0)
        # field-side pieces on junk J
        for ph in (0, 1):
            O.ghost_fold_phase(s, ph, 3); c.update_ghosts(3 + ph)
            assert np.array_equal(c.download_fields(), s.f)
        O.ghost_copy_axes(s, (6, 7, 8), 3); c.update_ghosts(1)
        assert np.array_equal(c.download_fields(), s.f)
        hp = (0.5 * k.px, 0.5 * k.py, 0.5 * k.pz)
        O.advance_b_stencil(s, *hp); c.advance_b_stencil(*hp)
        O.ghost_copy_axes(s, (3, 4, 5), 3); c.update_ghosts(2)
        O.advance_e_stencil(s, k.px, k.py, k.pz, k.dt_eps0); c.advance_e_stencil(k.px, k.py, k.pz, k.dt_eps0)
        assert np.array_equal(c.download_fields(), s.f)
        # push without the z wrap
        O.load_interpolator(s); c.load_interpolator_array()
        O.clear_accumulator(s); c.clear_accumulator_array()
        O.push(s, k, periodic=3); c.push(to_k(k))
        p = c.download_particles()
        for n in PARTICLE_NAMES:
            assert np.array_equal(p[n], s.p[n]), n
        assert acc_close(c.download_accumulators(), s.acc, prec)
        iz = s.p["cell"] // plane
        n_lo, n_hi = int((iz == 0).sum()), int((iz == nz + 1).sum())
        assert n_lo > 5 and n_hi > 5
        # extraction
        rb = np.dtype(PREC[prec]).itemsize
        cap = max(n_lo, n_hi) + 7
        lo = torch.zeros(cap * (7 * rb + 4), dtype=torch.uint8, device="cuda")
        hi = torch.zeros_like(lo)
        with pytest.raises(cp().CpicError) as err:          # send buffers too small: loud, store untouched
            c.extract_z_leavers(lo.data_ptr(), hi.data_ptr(), 2, 0, 0)
        assert err.value.code == -4 and c.num_particles == s.np
        got = c.extract_z_leavers(lo.data_ptr(), hi.data_ptr(), cap, 3 * plane, -nz * plane)
        assert got == (n_lo, n_hi)
        assert c.num_particles == s.np - n_lo - n_hi
        rest = c.download_particles()

        def unpack(buf, n):
            b = buf.cpu().numpy()
            d = {m: b[j * cap * rb: j * cap * rb + n * rb].view(PREC[prec]).copy() for j, m in enumerate(PARTICLE_NAMES[:7])}
            d["cell"] = b[7 * cap * rb: 7 * cap * rb + n * 4].view(np.int32).copy()
            return d
        for buf, n, sel, rebase in ((lo, n_lo, iz == 0, 3 * plane), (hi, n_hi, iz == nz + 1, -nz * plane)):
            d = unpack(buf, n)
            want = {m: s.p[m][sel] for m in PARTICLE_NAMES}
            want["cell"] = want["cell"] + rebase
            a, b = canonical_order(d), canonical_order(want)
            for m in PARTICLE_NAMES:
                assert np.array_equal(d[m][a], want[m][b]), m
        keep = (iz != 0) & (iz != nz + 1)
        want = {m: s.p[m][keep] for m in PARTICLE_NAMES}
        a, b = canonical_order(rest), canonical_order(want)
        for m in PARTICLE_NAMES:
            assert np.array_equal(rest[m][a], want[m][b]), m
        # append brings them back
        c.append_particles_device(lo.data_ptr(), cap, n_lo)
        c.append_particles_device(hi.data_ptr(), cap, n_hi)
        assert c.num_particles == s.np


def test_slab_stepper_single_rank_equals_fused_step():
    """world_size 1 run of the slab stepper (exchanges degenerate to local copies) against
    cpic_step on the same state: the choreography through GpuEngine reproduces the fused step.
    Fields to 2e-5 of scale (the ghost-plane accumulator rows are added as a lump), cells exact."""
    import torch
    from cabanapic_b200.dist import GpuEngine, SlabStepper
    m = cp()
    nx, ny, nz, prec = 6, 5, 7, "f32"
    s = random_state(nx, ny, nz, nppc=20, prec=prec, seed=3)
    k = to_k(consts_for(nx, ny, nz, prec))
    with make_ctx(s) as c:
        c.step(k, 5, 0, False)
        p_ref, f_ref = c.download_particles(), c.download_fields()
    e = GpuEngine(nx, ny, nz, s.np + 100, real=np.float32, z_periodic=False)
    try:
        e.ctx.upload_particles(s.p)
        e.ctx.upload_fields(s.f)
        st = SlabStepper(e, k, 0, 1, nz, nz, send_capacity=s.np)
        for _ in range(5):
            st.step()
        p, f = e.ctx.download_particles(), e.ctx.download_fields()
        assert st.migrated[0] > 0 and st.migrated[1] > 0
    finally:
        e.close()
    a, b = canonical_order(p), canonical_order(p_ref)
    assert np.mean(p["cell"][a] == p_ref["cell"][b]) > 0.999
    scale = np.abs(f_ref).max(axis=1, keepdims=True) + 1e-30
    interior = np.zeros((nz + 2, ny + 2, nx + 2), bool); interior[1:-1, 1:-1, 1:-1] = True
    assert (np.abs(f - f_ref) / scale)[:6, interior.ravel()].max() < 1e-4


def test_slab_stepper_fused_reorder_keeps_histogram_through_migration():
    """The slab stepper with the reordering push (what bench.py runs on N GPUs): extraction of the
    ghost-plane particles and the appended arrivals must keep the cell histogram of the last push
    exact, or the next push's cell segments would overflow into each other.  Same trajectories as the
    in-place stepper: the particle multisets agree bit for bit after 6 steps (strict mode), and the
    store ends up cell-ordered up to one step of drift."""
    from cabanapic_b200.dist import GpuEngine, SlabStepper
    nx, ny, nz, prec = 6, 5, 7, "f32"
    s = random_state(nx, ny, nz, nppc=20, prec=prec, seed=5)
    k = to_k(consts_for(nx, ny, nz, prec))
    out = []
    for fused in (False, True):
        e = GpuEngine(nx, ny, nz, s.np + 100, real=np.float32, z_periodic=False)
        try:
            e.ctx.upload_particles(s.p)
            e.ctx.upload_fields(s.f)
            st = SlabStepper(e, k, 0, 1, nz, nz, send_capacity=s.np)
            for _ in range(6):
                st.step(fused=fused)
            out.append((e.ctx.download_particles(), e.ctx.download_fields(), tuple(st.migrated)))
        finally:
            e.close()
    (p0, f0, m0), (p1, f1, m1) = out
    assert m0 == m1 and m0[0] > 0 and m0[1] > 0
    assert len(p0["cell"]) == len(p1["cell"]) == s.np
    a, b = canonical_order(p0), canonical_order(p1)
    # identical per-particle arithmetic; the fields differ by summation order only, which can flip a
    # rare borderline crossing after several steps
    assert np.mean(p0["cell"][a] == p1["cell"][b]) > 0.999
    scale = np.abs(f0).max(axis=1, keepdims=True) + 1e-30
    assert (np.abs(f1 - f0) / scale)[:6].max() < 1e-4


@pytest.mark.parametrize("reorder", [False, True])
def test_block_private_accumulator_equals_global_reductions(reorder, monkeypatch):
    """Small grids (<= 1024 cells incl. ghosts) deposit into a block-private shared-memory accumulator and
    histogram that join the global ones when the block retires (BASELINE configs[1]: 1e8 particles on 32 cells
    would otherwise serialise on 32 accumulator rows).  Against the global-reduction path on the same state:
    particles bit-identical, accumulators to summation order, and the next sort / reordering push (which consume
    the histogram) still produce a valid cell-ordered store."""
    s = random_state(5, 4, 3, nppc=300, prec="f32", seed=31)          # 7*6*5 = 210 cells, 18000 particles
    k = to_k(consts_for(5, 4, 3, "f32"))
    out = []
    for priv in ("0", "1"):
        monkeypatch.setenv("CPIC_PUSH2_PRIV", priv)
        with make_ctx(s) as c:
            c.sort_particles()
            for _ in range(2):
                c.load_interpolator_array(); c.clear_accumulator_array()
                c.push_reorder(k) if reorder else c.push(k)
            acc = c.download_accumulators()
            c.load_interpolator_array(); c.clear_accumulator_array()
            c.push_reorder(k)                                           # consumes the histogram of the last push
            p = c.download_particles()
            assert c.num_particles == s.np
            out.append((p, acc))
    (p0, a0), (p1, a1) = out
    o0, o1 = canonical_order(p0), canonical_order(p1)
    for n in PARTICLE_NAMES:
        assert np.array_equal(p0[n][o0], p1[n][o1]), n
    assert acc_close(a1, a0, "f32")
    assert np.abs(a0).max() > 0


def test_slab_runner_cuda_graph_replay_equals_eager_steps():
    """What bench.py runs on N GPUs, on one rank with the z exchange kept (the slab is its own neighbour): 8 steps
    replayed from a CUDA graph of two fused slab steps (device-counted migration, no host synchronisation) against
    the same 8 steps launched eagerly.  Same particle count, same migration totals up to the rare borderline
    crossing, cells equal for > 99.9 % of the particles, fields to 1e-4 of scale (atomic order differs per run)."""
    import torch
    from cabanapic_b200 import decks
    from cabanapic_b200.dist import SlabBench
    m = cp()
    d = decks.uniform_plasma(12, 10, 8, 16)
    k, _, we = d.consts()
    out = []
    for use_graph in (False, True):
        r = SlabBench(d, k, we, 0, 1, 0, m.FP_STRICT)
        r.open_z = True
        r.setup()
        try:
            r.step(2, -1)                         # warm-up: first extraction scans, then the pushes list their leavers
            if use_graph:
                r.prepare_timed(-1)
                assert r.graph is not None, "graph capture failed"
            r.step(8, -1)
            with r._on_stream():
                n = r.eng.ctx.num_particles
                out.append((r.eng.ctx.download_particles(), r.eng.ctx.download_fields(), n, list(r.stepper.migrated)))
        finally:
            r.close()
    (p0, f0, n0, m0), (p1, f1, n1, m1) = out
    assert n0 == n1 == d.num_particles
    assert all(abs(a - b) <= 2 + 0.002 * a for a, b in zip(m0, m1)) and m0[0] > 0 and m0[1] > 0
    a, b = canonical_order(p0), canonical_order(p1)
    assert np.mean(p0["cell"][a] == p1["cell"][b]) > 0.999
    scale = np.abs(f0).max(axis=1, keepdims=True) + 1e-30
    assert (np.abs(f1 - f0) / scale)[:6].max() < 1e-4


def test_slab_async_migration_counts_and_overflow():
    """cpic_slab_extract_async / cpic_slab_append_async (every count on the device, no host round trip) against
    the synchronising cpic_extract_z_leavers on the same states: same leaver counts, same particles left behind,
    same store after the arrivals are appended; a send-buffer overflow surfaces as CPIC_E_CAPACITY at the next
    call that needs the host's particle count."""
    import torch
    from cabanapic_b200.dist import GpuEngine
    m = cp()
    nx, ny, nz, prec = 6, 5, 4, "f32"
    s = random_state(nx, ny, nz, nppc=30, prec=prec, seed=9)
    k = to_k(consts_for(nx, ny, nz, prec))
    plane = (nx + 2) * (ny + 2)
    cap = s.np
    res = []
    for mode in ("sync", "async"):
        e = GpuEngine(nx, ny, nz, s.np + 100, real=np.float32, z_periodic=False)
        try:
            c = e.ctx
            c.upload_particles(s.p); c.upload_fields(s.f)
            lo, hi = e.alloc_bytes(cap * 32), e.alloc_bytes(cap * 32)
            cnt = torch.zeros(2, dtype=torch.int64, device=lo.device)
            counts = []
            for it in range(3):
                c.load_interpolator_array(); c.clear_accumulator_array(); c.push_reorder(k)
                if mode == "sync":
                    n_lo, n_hi = c.extract_z_leavers(lo.data_ptr(), hi.data_ptr(), cap, nz * plane, -nz * plane)
                    c.append_particles_device(lo.data_ptr(), cap, n_lo)      # periodic with itself: they come back
                    c.append_particles_device(hi.data_ptr(), cap, n_hi)
                else:
                    c.slab_extract_async(lo.data_ptr(), hi.data_ptr(), cap, cnt.data_ptr(), nz * plane, -nz * plane)
                    c.slab_append_async(lo.data_ptr(), cap, cnt[0:1].data_ptr())
                    c.slab_append_async(hi.data_ptr(), cap, cnt[1:2].data_ptr())
                    n_lo, n_hi = cnt.tolist()
                counts.append((n_lo, n_hi))
            assert c.num_particles == s.np
            res.append((counts, c.download_particles()))
            if mode == "async":          # overflow of a 2-particle send buffer on a device-counted extraction
                c.load_interpolator_array(); c.clear_accumulator_array(); c.push_reorder(k)
                c.slab_extract_async(lo.data_ptr(), hi.data_ptr(), 2, cnt.data_ptr(), nz * plane, -nz * plane)
                with pytest.raises(m.CpicError) as err:
                    c.num_particles
                assert err.value.code == -4
        finally:
            e.close()
    (c0, p0), (c1, p1) = res
    assert c0 == c1 and all(a > 0 and b > 0 for a, b in c0)
    a, b = canonical_order(p0), canonical_order(p1)
    for n in PARTICLE_NAMES:
        assert np.array_equal(p0[n][a], p1[n][b]), n


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("interval", [1, 2, 3, -1])
def test_sorted_steps_match_unsorted_oracle(prec, interval):
    """cpic_step with the periodic sort switched on (the push hands the histogram of the new cells to
    the following counting sort; the float path uses the pair-wise scatter kernel) against the oracle,
    which never reorders.  Sorting must be physics-neutral: after 7 steps the particle multisets agree
    -- cells bit-exact, state to the accumulated summation-order noise -- and the particles are
    cell-sorted whenever the last step started with a sort."""
    nx, ny, nz = 6, 5, 4
    s = random_state(nx, ny, nz, nppc=37, prec=prec, seed=8)       # odd count per cell: pairs straddle cells
    k = consts_for(nx, ny, nz, prec)
    O = Restatement(prec)
    nsteps = 7
    with make_ctx(s) as c:
        c.step(to_k(k), nsteps, sort_interval=interval, energies=False)
        p, f = c.download_particles(), c.download_fields()
        c.sort_particles()                                          # consumes a pending histogram if there is one
        ps = c.download_particles()
    O.step(s, k, 0, nsteps)
    assert len(p["cell"]) == s.np
    a, b = canonical_order(p), canonical_order(s.p)
    tol = 5e-5 if prec == "f32" else 1e-11
    assert np.mean(p["cell"][a] == s.p["cell"][b]) > 0.999
    assert np.array_equal(np.bincount(p["cell"], minlength=s.nc) > 0, np.bincount(s.p["cell"], minlength=s.nc) > 0)
    same = p["cell"][a] == s.p["cell"][b]
    for n in ("dx", "dy", "dz", "ux", "uy", "uz"):
        assert np.abs(p[n][a][same] - s.p[n][b][same]).max() < tol, n
    scale = np.abs(s.f).max(axis=1, keepdims=True) + 1e-30
    assert (np.abs(f - s.f) / scale).max() < (2e-4 if prec == "f32" else 1e-10)
    assert np.all(np.diff(ps["cell"]) >= 0)
    o1, o2 = canonical_order(p), canonical_order(ps)
    for n in PARTICLE_NAMES:
        assert np.array_equal(p[n][o1], ps[n][o2]), n


@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("fp", ["strict", "contract"])
@pytest.mark.parametrize("grid", GRIDS)
def test_push_reorder_teacher_forced(grid, fp, shuffle):
    """cpic_push_reorder (push + cell ordering in one pass, float): three consecutive calls, each from the
    oracle's inputs.  Bars: the particle multiset after every call is bit-identical to the oracle's push
    (strict mode; contract: 2e-6), mover / crossing counts exact, accumulators to summation order, and the
    store comes back ordered by the cell each particle occupied BEFORE that call (the segments come from the
    previous call's histogram from the second call on)."""
    nx, ny, nz = grid
    s = random_state(nx, ny, nz, nppc=37, prec="f32", seed=5)
    if shuffle:                                   # arbitrary input order: no cell runs for the claims to merge
        perm = np.random.default_rng(1).permutation(s.np)
        for n in PARTICLE_NAMES:
            s.p[n][:] = s.p[n][perm]
    k = consts_for(nx, ny, nz, "f32")
    O = Restatement("f32")
    m = cp()
    with make_ctx(s, fp_mode=m.FP_CONTRACT if fp == "contract" else m.FP_STRICT) as c:
        c.enable_push_stats(True)
        for call in range(3):
            old_cell = s.p["cell"].copy()
            O.load_interpolator(s); c.load_interpolator_array()
            O.clear_accumulator(s); c.clear_accumulator_array()
            movers, crossings = O.push(s, k)
            c.push_reorder(to_k(k))
            p = c.download_particles()
            assert len(p["cell"]) == s.np
            a, b = canonical_order(p), canonical_order(s.p)
            if fp == "strict":
                for n in PARTICLE_NAMES:
                    assert np.array_equal(p[n][a], s.p[n][b]), (call, n)
                st = c.push_stats()
                assert st["movers"] == movers and st["crossings"] == crossings and movers > 0
                # order of the store = order of the cells before the call
                before = np.empty(s.np, dtype=np.int64)
                before[a] = old_cell[b]
                assert np.all(np.diff(before) >= 0), call
            else:
                assert np.array_equal(np.sort(p["w"]), np.sort(s.p["w"]))
                assert np.mean(p["cell"][a] == s.p["cell"][b]) > 0.995
            assert acc_close(c.download_accumulators(), s.acc, "f32")
            if fp == "contract":                  # keep the oracle and the device on the same inputs
                c.upload_particles(s.p)


def test_push_reorder_edge_cases_and_fallbacks():
    """Odd / tiny / empty particle counts, a context without the second buffer, double precision
    (= cpic_sort_particles + cpic_push), and mixing with the in-place push and the sort."""
    m = cp()
    k = consts_for(3, 3, 2, "f32")
    O = Restatement("f32")
    for npart in (0, 1, 2, 63, 64, 65, 129):
        s = random_state(3, 3, 2, nppc=8, prec="f32", seed=3)
        keep = np.random.default_rng(npart).permutation(s.np)[:npart]
        from oracle.api import State
        t = State(3, 3, 2, 1, npart, "f32")
        for n in PARTICLE_NAMES:
            t.p[n][:] = s.p[n][keep]
        t.f[:] = s.f
        with make_ctx(t) as c:
            O.load_interpolator(t); c.load_interpolator_array()
            O.clear_accumulator(t); c.clear_accumulator_array()
            O.push(t, k); c.push_reorder(to_k(k))
            c.sort_particles()                    # the pending histogram feeds a plain sort as well
            p = c.download_particles()
            a, b = canonical_order(p), canonical_order(t.p)
            for n in PARTICLE_NAMES:
                assert np.array_equal(p[n][a], t.p[n][b]), (npart, n)
            assert np.all(np.diff(p["cell"]) >= 0)
            assert acc_close(c.download_accumulators(), t.acc, "f32")
            # and an in-place push afterwards still works on the reordered store
            O.load_interpolator(t); c.load_interpolator_array()
            O.clear_accumulator(t); c.clear_accumulator_array()
            O.push(t, k); c.push(to_k(k))
            O.load_interpolator(t); c.load_interpolator_array()
            O.clear_accumulator(t); c.clear_accumulator_array()
            O.push(t, k); c.push_reorder(to_k(k))
            p = c.download_particles()
            a, b = canonical_order(p), canonical_order(t.p)
            for n in PARTICLE_NAMES:
                assert np.array_equal(p[n][a], t.p[n][b]), (npart, n)
    s = random_state(3, 3, 2, nppc=8, prec="f32", seed=3)
    with make_ctx(s, enable_sort=False) as c:
        with pytest.raises(m.CpicError):
            c.push_reorder(to_k(k))
    s = random_state(3, 3, 2, nppc=8, prec="f64", seed=3)
    k = consts_for(3, 3, 2, "f64")
    O = Restatement("f64")
    with make_ctx(s) as c:
        O.load_interpolator(s); c.load_interpolator_array()
        O.clear_accumulator(s); c.clear_accumulator_array()
        O.push(s, k); c.push_reorder(to_k(k))
        p = c.download_particles()
        a, b = canonical_order(p), canonical_order(s.p)
        for n in PARTICLE_NAMES:
            assert np.array_equal(p[n][a], s.p[n][b]), n


@pytest.mark.gpu
@pytest.mark.parametrize("grid", [(6, 5, 4), (14, 12, 10)])      # few cells (k_push2, block-private) / k_push3
@pytest.mark.parametrize("fused", [False, True])
def test_two_species_share_the_accumulator(grid, fused):
    """two particle lists with their own charge and mass (decks/vpic/2stream-em0.cxx:207-208) pushed
    into ONE accumulator.  cpic_create_species + cpic_step_species against the oracle running its push twice per step
    on two particle sets that share the field / interpolator / accumulator arrays: per-species particle state
    bit-exact in strict mode while the fields agree to summation order (teacher-forced each step)."""
    from oracle.api import Consts as OConsts
    nx, ny, nz = grid
    se = random_state(nx, ny, nz, nppc=20, prec="f32", seed=31)
    si = random_state(nx, ny, nz, nppc=9, prec="f32", seed=32, uth=0.05)
    si.f, si.interp, si.acc = se.f, se.interp, se.acc                  # the ions see and feed the same arrays
    ke = consts_for(nx, ny, nz, "f32", qdt_2mc=-0.05)
    d = ke.to_dict()
    d.update(qsp=1.0, qdt_2mc=0.05 / 4.0)                              # ions: opposite charge, four times the mass
    ki = OConsts(**d)
    O = Restatement("f32")
    m = cp()
    hp = (0.5 * ke.px, 0.5 * ke.py, 0.5 * ke.pz)
    with make_ctx(se) as c:
        sp = c.create_species(si.np)
        try:
            sp.upload_particles(si.p)
            for step in range(3):
                O.load_interpolator(se); O.clear_accumulator(se)
                O.push(se, ke); O.push(si, ki)
                O.unload_accumulator(se, ke)
                O.advance_b(se, *hp); O.advance_e(se, ke.px, ke.py, ke.pz, ke.dt_eps0); O.advance_b(se, *hp)
                c.step_species([c, sp], [to_k(ke), to_k(ki)], 1, m.SORT_FUSED if fused else 0)
                for ctx_, s_ in ((c, se), (sp, si)):
                    p = ctx_.download_particles()
                    assert len(p["cell"]) == s_.np
                    a, b = canonical_order(p), canonical_order(s_.p)
                    for n in PARTICLE_NAMES:
                        assert np.array_equal(p[n][a], s_.p[n][b]), (step, n)
                f = c.download_fields()
                scale = np.abs(se.f).max(axis=1, keepdims=True) + 1e-30
                assert (np.abs(f - se.f) / scale).max() < 2e-4
                c.upload_fields(se.f)                                   # teacher-forced: same inputs for the next step
        finally:
            sp.close()


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("grid,fused", [((6, 5, 4), False), ((6, 5, 4), True), ((14, 12, 10), True)])
def test_reflect_boundary_matches_oracle(grid, fused, prec):
    """SURVEY 8f.3 -- Boundary::Reflect: reflecting particle walls (VPIC's block the reference keeps in comments,
    src/move_p.h:298-324) + a perfectly conducting box (E_tang = 0, src/grid.h:4-17).  The reference exit(1)s here, so
    parity is pinned by the oracle restatement only (oracle/cpic_oracle.c: move_particle / orc_pec_walls).  Teacher-
    forced steps: particle state bit-exact in strict mode (cells, positions on the wall, reversed momenta), fields to
    summation order; no particle ever leaves the interior; every drain variant (k_push / k_push2 block-private /
    k_push3) is exercised through the grids and precisions."""
    if fused and prec == "f64":
        pytest.skip("the reordering push is float")
    nx, ny, nz = grid
    s = random_state(nx, ny, nz, nppc=16, prec=prec, seed=41)
    k = consts_for(nx, ny, nz, prec)
    O = Restatement(prec)
    O.pec_walls(s)
    m = cp()
    R = PREC[prec]
    hp = (float(R(0.5) * R(k.px)), float(R(0.5) * R(k.py)), float(R(0.5) * R(k.pz)))
    with make_ctx(s, boundary=m.BOUNDARY_REFLECT) as c:
        c.enable_push_stats(True)
        reflected = 0
        for step in range(4):
            ux0 = s.p["ux"].copy()
            O.load_interpolator(s); c.load_interpolator_array()
            O.clear_accumulator(s); c.clear_accumulator_array()
            movers, crossings = O.push(s, k, periodic=7 << 4)
            (c.push_reorder if fused else c.push)(to_k(k))
            st = c.push_stats()
            assert st["movers"] == movers and st["crossings"] == crossings
            p = c.download_particles()
            a, b = canonical_order(p), canonical_order(s.p)
            for n in PARTICLE_NAMES:
                assert np.array_equal(p[n][a], s.p[n][b]), (step, n)
            assert acc_close(c.download_accumulators(), s.acc, prec)
            O.unload_accumulator(s, k); c.unload_accumulator_array(to_k(k))
            O.advance_b_stencil(s, *hp); c.advance_b(*hp)
            O.advance_e_stencil(s, k.px, k.py, k.pz, k.dt_eps0); O.pec_walls(s); c.advance_e(k.px, k.py, k.pz, k.dt_eps0)
            O.advance_b_stencil(s, *hp); c.advance_b(*hp)
            f = c.download_fields()
            scale = np.abs(s.f).max(axis=1, keepdims=True) + 1e-30
            assert (np.abs(f - s.f) / scale).max() < (2e-4 if prec == "f32" else 1e-11)
            c.upload_fields(s.f)
            ix = s.p["cell"] % (nx + 2); iy = (s.p["cell"] // (nx + 2)) % (ny + 2); iz = s.p["cell"] // ((nx + 2) * (ny + 2))
            assert ix.min() >= 1 and ix.max() <= nx and iy.min() >= 1 and iy.max() <= ny and iz.min() >= 1 and iz.max() <= nz
            reflected += int(np.sum(np.abs(s.p["dx"]) == 1.0))
        assert reflected >= 0
        d = c.state_digest()
        assert d["particles"] == s.np and d["cells_not_interior"] == 0 and d["offsets_out_of_range"] == 0


# ----------------------------------------------------------------------------- BASELINE configs[0] (C1) at its full length
def _two_stream_short_oracle(prec, real):
    from cabanapic_b200 import decks
    d = decks.two_stream_short(real, "x")
    k, _, _ = d.consts()
    p, f = d.initial_particles(), d.initial_fields()
    s = State(d.nx, d.ny, d.nz, 1, len(p["cell"]), prec)
    for n in PARTICLE_NAMES:
        s.p[n][:] = p[n]
    s.f[:] = f
    ok = OConsts.from_dict({n: getattr(k, n) for n, _ in OConsts._fields_})
    return d, np.array(Restatement(prec).step(s, ok, 0, d.num_steps, energies=True))


def test_two_stream_short_full_history_double():
    """decks/2stream-short.cxx:27-56 (32 cells, 3200 particles, 3000 steps; the x-oriented repair of SURVEY F1) in
    double, GPU vs the oracle over the deck's whole length.  The run is chaotic after saturation (~step 700): a mere
    permutation of the particles moves the SERIAL oracle's double history by 1e-9 up to step 1500, 1e-5 up to 2000 and
    23 % by 3000 (measured when this test was written).  Bars: 1e-6 up to step 1500, 1e-3 up to 2000, then the mean
    field energy of the last 500 steps within 15 %."""
    m = cp()
    d, want = _two_stream_short_oracle("f64", np.float64)
    sim = m.Simulation(d)
    en = sim.run(d.num_steps, energies=True)
    sim.close()
    assert en.shape == want.shape == (3000, 2)
    rel = np.abs(en[:, 0] - want[:, 0]) / want[:, 0]
    assert rel[:1500].max() < 1e-6, rel[:1500].max()
    assert rel[:2000].max() < 1e-3, rel[:2000].max()
    assert abs(en[2500:, 0].mean() / want[2500:, 0].mean() - 1) < 0.15
    assert want[:, 0].max() > 1e12 * want[0, 0]          # the instability really grew out of the seed (1e-14 -> 0.24)


def test_two_stream_short_full_history_float():
    """The same run in float (what the reference's default build computes).  Float histories decorrelate during the
    linear phase already (the serial oracle: 5 % by step 1000 under a permutation), so the physics is judged: growth
    out of the seed to the same saturation level (10 %) at the same time (30 steps of 3000), and the same late-time
    mean field energy (25 %)."""
    m = cp()
    d, want = _two_stream_short_oracle("f32", np.float32)
    sim = m.Simulation(d)
    en = sim.run(d.num_steps, energies=True)
    sim.close()
    assert abs(en[:, 0].max() / want[:, 0].max() - 1) < 0.10
    assert abs(int(en[:, 0].argmax()) - int(want[:, 0].argmax())) <= 30
    assert abs(en[2000:, 0].mean() / want[2000:, 0].mean() - 1) < 0.25
    # linear phase: the growth of the envelope (running maximum) between steps 200 and 600 agrees to 20 % in the exponent
    g = np.log(np.maximum.accumulate(en[:, 0])[600] / np.maximum.accumulate(en[:, 0])[200])
    w = np.log(np.maximum.accumulate(want[:, 0])[600] / np.maximum.accumulate(want[:, 0])[200])
    assert abs(g / w - 1) < 0.2, (g, w)


@pytest.mark.parametrize("reorder", [False, True])
def test_odd_count_ignores_the_padding_record(reorder):
    """with an odd particle count the last pair's B slot, rec[np], may hold anything.  Poison it with NaN /
    Inf (upload np+1 particles whose last one is non-finite, then lower the count) and push: the accumulators must be
    finite and equal to the oracle's, the np real particles bit-exact -- in place and through the reordering push."""
    m = cp()
    nx, ny, nz = 9, 7, 5                                   # > 1024 cells with ghosts: the float path is k_push3 when reordering
    s = random_state(nx, ny, nz, nppc=3, seed=5)
    n = s.np if s.np % 2 == 1 else s.np - 1                # an odd count
    t = State(nx, ny, nz, 1, n, "f32")
    for name in PARTICLE_NAMES:
        t.p[name][:] = s.p[name][:n]
    t.f[:] = s.f
    poisoned = {name: np.concatenate([t.p[name], t.p[name][:1]]) for name in PARTICLE_NAMES}
    for name, bad in (("dx", np.nan), ("dy", np.inf), ("ux", np.nan), ("uy", -np.inf), ("uz", 3e38), ("w", np.nan)):
        poisoned[name][-1] = bad
    kk = consts_for(nx, ny, nz)
    O = Restatement("f32")
    O.load_interpolator(t); O.clear_accumulator(t); O.push(t, kk)
    with m.Context(nx, ny, nz, 1, max_particles=n + 65, real=np.float32) as c:
        c.upload_particles(poisoned)
        c.set_num_particles(n)
        c.upload_fields(s.f)
        c.load_interpolator_array(); c.clear_accumulator_array()
        if reorder:
            c.push_reorder(to_k(kk))
        else:
            c.push(to_k(kk))
        acc = c.download_accumulators()
        p = c.download_particles()
    assert np.isfinite(acc).all()
    assert acc_close(acc, t.acc, "f32")
    og, oo = canonical_order(p), canonical_order(t.p)
    for name in PARTICLE_NAMES:
        assert np.array_equal(p[name][og], t.p[name][oo]), name


# ----------------------------------------------------------------------------- CPIC_DEPOSIT_ORDERED: the reference's summation order
@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("grid", [(6, 5, 4), (1, 32, 1), (17, 9, 5)])
def test_ordered_deposit_is_bitwise_over_whole_steps(grid, prec):
    """CPIC_DEPOSIT_ORDERED adds the streaks of every cell in (particle, streak) order -- the order of the reference's
    serial loop -- so in strict FP mode NOTHING differs from the oracle any more: the accumulators after one push, and
    particles, accumulators and all nine field members after several whole steps, bit for bit (float and double; the
    other deposit modes agree to summation order only: acc_close)."""
    m = cp()
    nx, ny, nz = grid
    kk = consts_for(nx, ny, nz, prec)
    O = Restatement(prec)
    s = random_state(nx, ny, nz, nppc=9, prec=prec, seed=13)
    with make_ctx(s, deposit_mode=m.DEPOSIT_ORDERED, enable_sort=False) as c:
        O.load_interpolator(s); O.clear_accumulator(s); O.push(s, kk)
        c.load_interpolator_array(); c.clear_accumulator_array(); c.push(to_k(kk))
        assert np.array_equal(c.download_accumulators(), s.acc)
        p = c.download_particles()
        for n in PARTICLE_NAMES:
            assert np.array_equal(p[n], s.p[n]), n
    s = random_state(nx, ny, nz, nppc=9, prec=prec, seed=14)
    with make_ctx(s, deposit_mode=m.DEPOSIT_ORDERED, enable_sort=False) as c:
        O.step(s, kk, 0, 6)
        c.step(to_k(kk), 6, sort_interval=0, energies=False)
        p = c.download_particles()
        for n in PARTICLE_NAMES:
            assert np.array_equal(p[n], s.p[n]), n
        assert np.array_equal(c.download_fields(), s.f)
        assert np.array_equal(c.download_accumulators(), s.acc)


def test_ordered_deposit_history_is_bitwise_float():
    """The float run of the reference's regression deck (2stream-em, 1x32x1) -- chaotic after saturation under any other
    summation order -- is reproduced bit for bit over 1500 steps with the ordered deposit: identical particles at the
    end, energy history equal to the float rounding of the diagnostic itself."""
    from cabanapic_b200 import decks
    m = cp()
    d = decks.two_stream_em(np.float32)
    k, _, _ = d.consts()
    p, f = d.initial_particles(), d.initial_fields()
    s = State(d.nx, d.ny, d.nz, 1, len(p["cell"]), "f32")
    for n in PARTICLE_NAMES:
        s.p[n][:] = p[n]
    s.f[:] = f
    ok = OConsts.from_dict({n: getattr(k, n) for n, _ in OConsts._fields_})
    want = np.array(Restatement("f32").step(s, ok, 0, 1500, energies=True))
    sim = m.Simulation(d, deposit_mode=m.DEPOSIT_ORDERED, enable_sort=False)
    en = sim.run(1500, sort_interval=0, energies=True)
    got = sim.particles()
    sim.close()
    # the state is bit-identical at every step; the energy DIAGNOSTIC is summed in double here and in the working
    # precision by the reference (src/fields.h:722-763), hence the 2e-5 of a..19
    assert np.allclose(en, want, rtol=2e-5, atol=0)
    for n in PARTICLE_NAMES:
        assert np.array_equal(got[n], s.p[n]), n

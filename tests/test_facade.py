"""The C++ host facade (include/cabanapic/src/*.h): the reference's decks -- and the reference's own
main() -- compiled UNMODIFIED against our headers, running on the GPU through the C ABI.

CPU part: the facade builds for every deck, and a binary without a GPU fails loudly (no fallback).
GPU part: the reference's smoke test (tests/decks: custom_init must exit 0) with the reference's
own example.cpp; energy histories of our driver against the committed fixtures generated from the
reference build; the reference's own regression test (tests/energy_comparison: 6000 steps of
2stream-em in double, judged by the reference's finalizer against its gold file)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "examples", "build")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
HAVE_REF = os.path.isdir("/root/reference/decks")


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, stdout=subprocess.DEVNULL)


def _run(name, env=None, cwd=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([os.path.join(BUILD, name)], cwd=cwd, env=e, capture_output=True, text=True, timeout=timeout)


def test_facade_builds_for_every_deck():
    _build()
    want = ["cbnpic_default", "cbnpic_weibel_3d"]
    if HAVE_REF:      # the reference's decks and main(), compiled unmodified where they lie
        want += ["cbnpic_custom_init", "cbnpic_dioctron_3d", "cbnpic_2particle", "cbnpic_2stream-em",
                 "cbnpic_custom_init_es", "cbnpic_2stream-em_double", "ref_custom_init"]
    for w in want:
        assert os.path.exists(os.path.join(BUILD, w)), w


def test_facade_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _build()
    r = _run("cbnpic_default", cwd=tmp_path)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr


def _energies(path):
    return np.loadtxt(path, ndmin=2)


needs_bins = pytest.mark.skipif(not os.path.exists(os.path.join(BUILD, "ref_custom_init")),
                                reason="examples/build was not produced in the build container")


@pytest.mark.gpu
@needs_bins
def test_reference_main_and_deck_unmodified_smoke(tmp_path):
    """tests/decks/CMakeLists.txt:3-15 of the reference: custom_init must run to completion (exit 0).
    Here: the reference's example/example.cpp + decks/custom_init.cxx, both unmodified, on the facade."""
    r = _run("ref_custom_init", cwd=tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    en = _energies(tmp_path / "energies.txt")
    z = np.load(os.path.join(GOLDEN, "state_custom_init_f32.npz"))
    assert en.shape[0] == 30
    ref = z["energies"]
    # noise-seeded two-stream energies: judged against the history's scale (see test_gpu_parity.py)
    assert np.abs(en[:, 2] - ref[:, 0]).max() < 2e-3 * ref[:, 0].max()


@pytest.mark.gpu
@needs_bins
@pytest.mark.parametrize("binary,fixture,steps", [("cbnpic_custom_init", "custom_init_f32", 30),
                                                  ("cbnpic_dioctron_3d", "dioctron_3d_f32", 10),
                                                  ("cbnpic_2particle", "2particle_f32", 200)])
def test_driver_energy_history_vs_reference_fixture(tmp_path, binary, fixture, steps):
    """Our driver + the reference's deck (unmodified) vs the energies the reference build produced
    (tests/golden/state_*.npz).  energies.txt carries 6 significant digits and the float deposit sums
    in another order; the noise-seeded histories are judged against their own scale: 2e-3 of the maximum."""
    r = _run(binary, env={"CPIC_STEPS": str(steps)}, cwd=tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    en = _energies(tmp_path / "energies.txt")
    ref = np.load(os.path.join(GOLDEN, f"state_{fixture}.npz"))["energies"]
    assert en.shape[0] == steps
    for col in (0, 1):
        assert np.abs(en[:, 2 + col] - ref[:steps, col]).max() <= 2e-3 * ref[:steps, col].max() + 1e-30, col


@pytest.mark.gpu
@needs_bins
def test_reference_regression_test_passes_on_gpu(tmp_path):
    """The reference's one real regression test (tests/energy_comparison): 6000 steps of the 1x32x1
    two-stream deck in double; the deck's own Custom_Finalizer compares energies.txt with the gold
    file (10 % on lines 3581..4880, tests/energy_comparison/2stream-em.cxx:23,45-70) and exits 1 on
    mismatch.  Deck and finalizer are the reference's files, unmodified."""
    r = _run("cbnpic_2stream-em_double", cwd=tmp_path, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    en = _energies(tmp_path / "energies.txt")
    assert en.shape == (6000, 4)
    gold = np.load(os.path.join(GOLDEN, "energies_gold_2stream-em.npz"))
    lines, g = gold["lines"], gold["f64"]
    assert (np.abs(en[lines, 2:4] - g) / np.minimum(en[lines, 2:4], g)).max() < 1e-4


@pytest.mark.gpu
@needs_bins
def test_es_solver_build_and_weibel_deck(tmp_path):
    """-DES_FIELD_SOLVER build of custom_init (the reference CI's only ES coverage is this exit code);
    and the new 3-D deck: magnetic field energy must grow out of the noise (Weibel instability)."""
    r = _run("cbnpic_custom_init_es", cwd=tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    assert _energies(tmp_path / "energies.txt").shape == (30, 3)
    w = tmp_path / "weibel"
    w.mkdir()
    r = _run("cbnpic_weibel_3d", env={"CPIC_WEIBEL_N": "16", "CPIC_WEIBEL_PPC": "32", "CPIC_STEPS": "150"}, cwd=w)
    assert r.returncode == 0, r.stderr[-2000:]
    en = _energies(w / "energies.txt")
    assert en[-1, 3] > 5 * en[4, 3] > 0        # measured: x13 at 16^3 x 32 ppc (thermal noise floor is high)

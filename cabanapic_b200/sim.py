"""Host-side mirror of the reference driver (example/example.cpp:44-290) over the C ABI.

``Simulation`` owns one ``Context`` and walks the reference's start-up and time loop:
derive parameters and constants, run the deck's initialisers on the host, upload, optional
uncenter, then the per-step sequence.  ``energies.txt`` is written in the reference's format
(src/fields.h:752-761); the per-step ``partloc``/``ex1d`` ASCII dumps are opt-in.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import Consts, Context
from .decks import Deck


class Simulation:
    def __init__(self, deck: Deck, solver=_lib.SOLVER_EM, device=0, fp_mode=_lib.FP_STRICT,
                 deposit_mode=_lib.DEPOSIT_AUTO, enable_sort=True, particles=None, fields=None,
                 capacity_factor=1.0):
        self.deck = deck
        self.consts, self.dxp, self.we = deck.consts()       # example.cpp:61-113,179-181
        self.solver = solver
        n = deck.num_particles if particles is None else len(particles["cell"])
        self.ctx = Context(deck.nx, deck.ny, deck.nz, deck.num_ghosts, max_particles=int(n * capacity_factor) + 64,
                           real=deck.real, solver=solver, device=device, fp_mode=fp_mode,
                           deposit_mode=deposit_mode, enable_sort=enable_sort)
        self.ctx.upload_particles(deck.initial_particles() if particles is None else particles)   # :121-124
        self.ctx.upload_fields(deck.initial_fields() if fields is None else fields)               # :144-168
        self.step_count = 0
        self.energy_log = []
        if deck.perform_uncenter:                                                                 # :204-213
            self.ctx.load_interpolator_array()
            self.ctx.uncenter_particles(self.consts.qdt_2mc)

    def close(self):
        self.ctx.close()

    def step_unfused(self):
        """One step through the individual reference-named calls (example.cpp:221-266)."""
        c, k = self.ctx, self.consts
        R = self.deck.real
        hx, hy, hz = float(R(0.5) * R(k.px)), float(R(0.5) * R(k.py)), float(R(0.5) * R(k.pz))
        c.load_interpolator_array()
        c.clear_accumulator_array()
        c.push(k)
        c.contribute()
        c.unload_accumulator_array(k)
        c.advance_b(hx, hy, hz)
        c.advance_e(k.px, k.py, k.pz, k.dt_eps0)
        c.advance_b(hx, hy, hz)
        self.step_count += 1

    def run(self, nsteps, sort_interval=0, energies=True):
        """nsteps fused on the device (cpic_step); returns the (nsteps,2) energy history."""
        en = self.ctx.step(self.consts, nsteps, sort_interval, energies)
        if energies:
            for s in range(nsteps):
                self.energy_log.append((self.step_count + s + 1, en[s, 0], en[s, 1]))
        self.step_count += nsteps
        return en

    def write_energies(self, path="energies.txt"):
        """`step time e_energy [b_energy]`, one line per step (src/fields.h:752-761)."""
        R = self.deck.real
        with open(path, "w") as fh:
            for step, e, b in self.energy_log:
                t = float(R(step) * R(self.consts.dt))
                if self.solver == _lib.SOLVER_EM:
                    fh.write(f"{step} {t:g} {float(R(e)):g} {float(R(b)):g}\n")
                else:
                    fh.write(f"{step} {t:g} {float(R(e)):g}\n")

    def particles(self):
        return self.ctx.download_particles()

    def fields(self):
        return self.ctx.download_fields()

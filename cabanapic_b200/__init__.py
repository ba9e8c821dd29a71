"""cabanapic_b200 -- B200-native implementation of CabanaPIC's per-timestep particle hot path.

The product is the C-ABI shared library built from ``csrc/`` (hand-written sm_100a CUDA,
declared in ``include/cabanapic_b200.h``) plus the C++ host facade in ``include/cabanapic/``.
The Python modules are plumbing for tests and the benchmark:

* ``_lib``   ctypes binding of the C ABI (``Context``);
* ``decks``  host-side mirror of the reference's deck parameters and initialisers;
* ``sim``    host-side mirror of the reference driver loop (``Simulation``);
* ``dist``   multi-GPU modes (replicated grid / z-slabs) over ``torch.distributed``.
"""
from ._lib import (BOUNDARY_PERIODIC, BOUNDARY_REFLECT, DEPOSIT_ATOMIC, DEPOSIT_ATOMIC_V4, DEPOSIT_AUTO, DEPOSIT_ORDERED,
                   DEPOSIT_WARP, FP_CONTRACT, FP_STRICT, SOLVER_EM, SOLVER_ES_1D, SORT_FUSED, Consts, Context, CpicError,
                   MGPU_AUTO, MGPU_REPLICATED, MGPU_SLAB, Mgpu, build, lib)
from .decks import Deck
from .sim import Simulation

__all__ = ["Context", "Consts", "CpicError", "Deck", "Simulation", "build", "lib", "SOLVER_EM", "SOLVER_ES_1D",
           "BOUNDARY_PERIODIC", "BOUNDARY_REFLECT", "FP_STRICT", "FP_CONTRACT", "DEPOSIT_AUTO", "DEPOSIT_ATOMIC",
           "DEPOSIT_ATOMIC_V4", "DEPOSIT_WARP", "DEPOSIT_ORDERED", "SORT_FUSED", "Mgpu", "MGPU_AUTO", "MGPU_REPLICATED", "MGPU_SLAB"]

"""ERI entry points of IntegralsModule / SpecialMatricesModule on the B200 backend.

computeTensorBlockElectronRepulsionIntegrals(sh1,sh2,sh3,sh4)  -- SpecialMatricesModule.jl:184-186 /
    LibInt2Module.jl:319-334: the dense block [mu,nu,la,si] of one shell quartet.
computeElectronRepulsionIntegral(...)                          -- IntegralsModule.jl:447-462 for shell components.
computeTensorElectronRepulsionIntegrals(shells)                -- SpecialMatricesModule.jl:198-207 (N^4, small N).
"""
from __future__ import annotations

import numpy as np

from .b200shells import B200Shells
from .shell import Shell


def computeTensorBlockElectronRepulsionIntegrals(sh1: Shell, sh2: Shell, sh3: Shell, sh4: Shell) -> np.ndarray:
    tmp = B200Shells([sh1, sh2, sh3, sh4])
    try:
        return tmp.eri_block(0, 1, 2, 3)
    finally:
        tmp.destroy()


def computeTensorElectronRepulsionIntegrals(shells) -> np.ndarray:
    own = not isinstance(shells, B200Shells)
    b = B200Shells(shells) if own else shells
    try:
        return b.eri_tensor()
    finally:
        if own:
            b.destroy()

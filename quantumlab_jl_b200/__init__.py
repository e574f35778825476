"""qlb200: B200-native Fock-build engine behind QuantumLab.jl's SCF hot path.

Importing the package does not touch the GPU; the first B200Shells(...) does (and fails loudly without one)."""
from .base import Ionization, LQuantumNumber, MQuantumNumber, MQuantumNumbers, Position  # noqa: F401
from .shell import Shell, computeBasisShells, nbf  # noqa: F401
from .basisset import BasisSet, loadBasisSet, readBasisSetTX93  # noqa: F401
from .geometry import Geometry, computeEnergyInteratomicRepulsion, readGeometryXYZ  # noqa: F401
from .initialguess import SADGuess, ZeroGuess  # noqa: F401

"""Basic value types of the hot path (mirror of BaseModule.jl).

Position (BaseModule.jl:5-9), LQuantumNumber (:80-86), MQuantumNumber (:46-50), the Cartesian
component order MQuantumNumbers (:110-122), numberMQNsCartesian (:93-96), distance (:33-36) and
doublefactorial (:130).  The component order is part of the index contract of every matrix the
library returns.
"""
from __future__ import annotations

import math
from typing import NamedTuple

LQN_EXPONENT = {"S": 0, "P": 1, "D": 2, "F": 3, "G": 4, "H": 5, "I": 6, "K": 7}
EXPONENT_LQN = {v: k for k, v in LQN_EXPONENT.items()}


class Position(NamedTuple):
    x: float
    y: float
    z: float

    def __sub__(self, o):  # type: ignore[override]
        return Position(self.x - o.x, self.y - o.y, self.z - o.z)


origin = Position(0.0, 0.0, 0.0)


class MQuantumNumber(NamedTuple):
    x: int
    y: int
    z: int


class LQuantumNumber:
    """LQuantumNumber("D") or LQuantumNumber(2)."""

    __slots__ = ("symbol", "exponent")

    def __init__(self, v):
        if isinstance(v, str):
            self.symbol, self.exponent = v, LQN_EXPONENT[v]
        else:
            self.symbol, self.exponent = EXPONENT_LQN[int(v)], int(v)

    def __eq__(self, o):
        return isinstance(o, LQuantumNumber) and o.exponent == self.exponent

    def __lt__(self, o):
        return self.exponent < o.exponent

    def __hash__(self):
        return hash(self.exponent)

    def __repr__(self):
        return f"LQuantumNumber({self.symbol!r})"


def _lexp(lqn) -> int:
    return lqn.exponent if isinstance(lqn, LQuantumNumber) else int(lqn)


def numberMQNsCartesian(lqn) -> int:
    l = _lexp(lqn)
    return ((l + 1) ** 2 + (l + 1)) // 2


def MQuantumNumbers(lqn) -> list[MQuantumNumber]:
    """x descending, then y descending: d = xx,xy,xz,yy,yz,zz (BaseModule.jl:110-122)."""
    m = _lexp(lqn)
    return [MQuantumNumber(x, y, m - x - y) for x in range(m, -1, -1) for y in range(m - x, -1, -1)]


def distance(p: Position, q: Position) -> float:
    return math.sqrt((p.x - q.x) ** 2 + (p.y - q.y) ** 2 + (p.z - q.z) ** 2)


def doublefactorial(n: int) -> int:
    r = 1
    while n >= 1:
        r *= n
        n -= 2
    return r


class Ionization(NamedTuple):
    charge: int

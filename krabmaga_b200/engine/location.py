"""Real2D / Int2D value types (src/engine/location.rs:9-41)."""
from typing import NamedTuple


class Real2D(NamedTuple):
    x: float
    y: float


class Int2D(NamedTuple):
    x: int
    y: int

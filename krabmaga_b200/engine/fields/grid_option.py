"""`GridOption` (src/engine/fields/grid_option.rs:3-10)."""
import enum


class GridOption(enum.IntEnum):
    READ = 0
    WRITE = 1
    READWRITE = 2

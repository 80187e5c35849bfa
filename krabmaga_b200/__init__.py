"""krabmaga_b200 — B200-native implementation of krABMaga's agent-step hot path.

Only what the path needs: the CUDA library behind include/krabgpu.h (csrc/), its ctypes binding
(_abi) and the host-side mirror of the reference interface for this path (engine/, flockers,
simulate).  There is no CPU fallback; importing works anywhere, device calls need a B200.
"""
from . import _abi
from ._abi import KgBoidsParams, KgError, KgLifeRule, KgOutOfBounds, boids_params, build, life_rule
from .batch import FlockerBatch
from .engine.agent import Agent
from .engine.fields.dense_number_grid_2d import DenseNumberGrid2D
from .engine.fields.dense_object_grid_2d import DenseGrid2D
from .engine.fields.sparse_object_grid_2d import SparseGrid2D
from .plots import PlotData, addplot, plot, plot_series
from .custom import FieldAgents, FieldModel
from .engine.fields.field import Field
from .engine.fields.field_2d import Field2D
from .engine.fields.grid_option import GridOption
from .engine.location import Int2D, Real2D
from .engine.schedule import Schedule
from .engine.state import State
from .explore import (ExploreMode, explore_distributed, explore_parallel, explore_sequential, field_names,
                      write_csv)
from .flockers import Flock, Flocker
from .simulate import simulate, simulate_explore, simulate_old

__all__ = ["Agent", "DenseGrid2D", "DenseNumberGrid2D", "ExploreMode", "Field", "Field2D", "FieldAgents", "FieldModel", "Flock", "Flocker",
           "FlockerBatch", "GridOption", "explore_distributed", "explore_parallel", "explore_sequential",
           "Int2D", "PlotData", "addplot", "plot", "plot_series", "KgBoidsParams", "KgError", "KgLifeRule", "KgOutOfBounds", "life_rule", "Real2D", "Schedule", "SparseGrid2D", "State",
           "boids_params", "build", "field_names", "simulate", "simulate_explore", "simulate_old",
           "write_csv"]

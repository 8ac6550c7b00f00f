"""ivp_b200 -- B200-native batched IVP solves behind the API of the Rust crate `ivp`.

Public surface (mirrors reference src/prelude.rs:35-43 plus the batched call):
    Method, Status, Direction, EventConfig, Options, Solution, BatchSolution, ConfigError,
    Problem, Context, solve_ivp_batch, solve_ivp
"""
from .types import (BatchSolution, ConfigError, Direction, EventConfig, InterpolationError, Method,  # noqa: F401
                    Options, Solution, Status)

__all__ = ["InterpolationError", "BatchSolution", "ConfigError", "Direction", "EventConfig", "Method", "Options", "Solution", "Status",
           "Problem", "Context", "solve_ivp_batch", "solve_ivp", "builtin"]


def __getattr__(name):
    # The CUDA binding is imported lazily so that the pure-host types work without libivpb.so;
    # touching any compute entry point loads the library and fails loudly if it is missing.
    if name in ("Problem", "Context", "solve_ivp_batch", "solve_ivp", "builtin", "PROBLEMS"):
        from . import api
        return getattr(api, name)
    raise AttributeError(name)

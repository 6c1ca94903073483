"""simplemcmc.jl_b200 -- B200-native executor for SimpleMCMC.jl's hot path (log-density + reverse-mode gradient,
many chains in lockstep).  Host side mirrors the reference's exported API; the work happens in
libsimplemcmc_cuda.so (csrc/), reached through the C ABI of include/simplemcmc.h."""
from .api import (MCMCRun, ModelFunction, generateModelFunction, simpleHMC, simpleNUTS, simpleRWM)  # noqa: F401
from .lowering import ModelError  # noqa: F401
from .solvers import SolverRun, simpleAGD, simpleNM  # noqa: F401

"""june_b200 -- B200-native replacement for JUNE's per-timestep infection hot path
(``Simulator.do_timestep``).  Python host, hand-written sm_100a CUDA kernels behind a C ABI
(``include/june_b200.h``).  See DESIGN.md / INTEGRATION.md."""
from ._lib import SimulatorError  # noqa: F401
from .disease import DiseaseConfig  # noqa: F401
from .interaction import Interaction  # noqa: F401
from .world import WorldSoA  # noqa: F401

__version__ = "0.1.0"

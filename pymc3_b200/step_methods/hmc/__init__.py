from .hmc import HamiltonianMC  # noqa: F401
from .nuts import NUTS  # noqa: F401

from .hmc import NUTS, HamiltonianMC  # noqa: F401
from .arraystep import Competence  # noqa: F401

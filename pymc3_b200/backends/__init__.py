from .base import BaseTrace, MultiTrace, merge_traces  # noqa: F401
from .ndarray import NDArray, load_trace, save_trace  # noqa: F401

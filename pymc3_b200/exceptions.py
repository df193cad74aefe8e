"""pymc3/exceptions.py:24-31 equivalents."""


class SamplingError(RuntimeError):
    pass


class ParallelSamplingError(Exception):
    """pymc3/parallel_sampling.py:64-70: raised with the failing chain's id."""

    def __init__(self, message, chain, warnings=None):
        super().__init__(message)
        self._chain = chain
        self._warnings = warnings if warnings is not None else []

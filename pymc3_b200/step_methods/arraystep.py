"""pymc3/step_methods/arraystep.py:27-44: how well a step method suits a variable."""
import enum


class Competence(enum.IntEnum):
    INCOMPATIBLE = 0
    COMPATIBLE = 1
    PREFERRED = 2
    IDEAL = 3

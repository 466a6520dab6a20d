"""B200-native batched implementation of the reference's `Environment::step` hot path.

Host-side mirror of the reference's public interface for that path (`World`, `Move`,
`Environment` -> `BatchedEnvironment`, the GA of src/main.rs -> `GeneticAlgorithm`) over the
C ABI in include/prl.h. Importing this package loads libprl.so and raises if it is missing.
"""
from . import _lib
from ._lib import PrlError
from .world import Move, ObjectAndTransform, World, WorldObject
from .batched import BatchedEnvironment
from .genetic import GaPopulation, GeneticAgent, GeneticAlgorithm
from .parallel import gather_generation, slice_bounds

_lib.lib()  # fail loudly at import time if the native library is absent

__all__ = ["World", "WorldObject", "ObjectAndTransform", "Move", "BatchedEnvironment", "GeneticAlgorithm", "GeneticAgent",
           "GaPopulation", "gather_generation", "slice_bounds", "PrlError"]

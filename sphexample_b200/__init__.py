"""sphexample_b200 — B200-native SPH inner loop behind SPHExample's SimulationLoop API."""
from .config import *  # noqa: F401,F403
from .config import make_params, next_output_time  # noqa: F401
from .preprocess import (AllocateDataStructures, LoadBoundaryNormals, LoadMDBCNormals,  # noqa: F401
                         LoadSpecificCSV, SimParticles, make_particles)

"""``sim_matrix`` under the module name the reference's trainers import it from
(``from model.model import sim_matrix``, trainer/trainer_dist.py:7)."""
from .loss import sim_matrix  # noqa: F401

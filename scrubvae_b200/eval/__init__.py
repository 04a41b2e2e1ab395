from . import eval  # noqa: A004
from .eval import generative_restrictiveness, r2_score

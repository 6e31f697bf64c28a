"""Trainers on the B200 hot path: ``CTRTrainer`` (ranking) and ``MatchTrainer`` (two-tower retrieval with in-batch negatives,
SURVEY.md §8 f3).  MTL / sequence-generation trainers are out of scope."""
from .ctr_trainer import CTRTrainer
from .match_trainer import MatchTrainer

__all__ = ["CTRTrainer", "MatchTrainer"]

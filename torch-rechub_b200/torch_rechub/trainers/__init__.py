"""Trainers on the B200 hot path.  ``CTRTrainer`` only; Match/MTL/Seq trainers are out of scope."""
from .ctr_trainer import CTRTrainer

__all__ = ["CTRTrainer"]

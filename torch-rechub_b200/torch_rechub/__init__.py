"""torch_rechub — the B200-native (sm_100a) engine behind torch-rechub's Python API.

Same import name and module paths as datawhalechina/torch-rechub v0.8.0 for the CTR hot path
(``basic.features``, ``basic.layers``, ``models.ranking.{DeepFM, DCN, DCNv2, DIN}``,
``trainers.CTRTrainer``, ``utils.data``); the engine itself lives in ``torch_rechub.b200``.
"""
__version__ = "0.8.0+b200"
__author__ = "rechub-b200"
__license__ = "MIT"
__url__ = "https://github.com/datawhalechina/torch-rechub"

from . import basic, models, trainers, utils  # noqa: E402,F401

__all__ = ["__version__", "__author__", "__license__", "__url__", "basic", "models", "trainers", "utils"]

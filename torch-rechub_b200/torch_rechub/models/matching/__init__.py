"""Matching models on the B200 hot path: DSSM (SURVEY.md §8 f3 — the two-tower consumer of the fused gather, trained with
in-batch negatives by ``MatchTrainer``) and its two plain variants that exercise the trainer's other loss modes:
FaceBookDSSM (pair-wise, BPR) and YoutubeDNN (list-wise, softmax).

The reference exports nine more retrieval models (``models/matching/__init__.py:1-13``).  They are outside this engine's
scope: the names stay importable and raise on construction with a pointer to upstream.
"""
__all__ = ['DSSM', 'FaceBookDSSM', 'YoutubeDNN', 'YoutubeSBC', 'MIND', 'GRU4Rec', 'NARM', 'SASRec', 'SINE', 'STAMP', 'ComirecDR', 'ComirecSA']

from .dssm import DSSM
from .dssm_facebook import FaceBookDSSM
from .youtube_dnn import YoutubeDNN


def _out_of_scope(name):

    class _OutOfScope(object):
        __doc__ = "%s is not part of the B200 hot-path engine (matching: DSSM, FaceBookDSSM, YoutubeDNN); use upstream torch-rechub for it." % name

        def __init__(self, *args, **kwargs):
            raise NotImplementedError(self.__doc__)

    _OutOfScope.__name__ = _OutOfScope.__qualname__ = name
    return _OutOfScope


for _name in __all__:
    if _name not in globals():
        globals()[_name] = _out_of_scope(_name)
del _name

"""Ranking models on the B200 hot path: DeepFM, DCN, DCNv2, DIN, plus WideDeep
(the plainest EmbeddingLayer consumer; the reference's ranking e2e test drives it).

The reference exports more ranking models (``models/ranking/__init__.py:1-14``).  They are outside this
engine's scope: their names stay importable so the reference's example scripts
(``examples/ranking/run_criteo.py:10``) import unchanged, but constructing one raises with a pointer to
upstream.  They would sit on the same ``EmbeddingLayer`` front end unchanged.
"""
__all__ = ['WideDeep', 'DeepFM', 'DCN', 'DCNv2', 'EDCN', 'AFM', 'FiBiNet', 'DeepFFM', 'BST', 'DIN', 'DIEN', 'FatDeepFFM', 'AutoInt']

from .dcn import DCN
from .dcn_v2 import DCNv2
from .deepfm import DeepFM
from .din import DIN
from .widedeep import WideDeep


def _out_of_scope(name):

    class _OutOfScope(object):
        __doc__ = "%s is not part of the B200 hot-path engine (DeepFM / DCN / DCNv2 / DIN); use upstream torch-rechub for it." % name

        def __init__(self, *args, **kwargs):
            raise NotImplementedError(self.__doc__)

    _OutOfScope.__name__ = _OutOfScope.__qualname__ = name
    return _OutOfScope


for _name in __all__:
    if _name not in globals():
        globals()[_name] = _out_of_scope(_name)
del _name

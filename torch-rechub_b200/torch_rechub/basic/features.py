"""Feature descriptors (mirror of reference ``torch_rechub/basic/features.py:5-87``).

A descriptor owns (lazily) the embedding table of its field: ``get_embedding_layer()`` builds it
once through the descriptor's initializer and caches it as ``.embed`` — so a descriptor re-used by
two models shares ONE table (reference ``features.py:36-39,68-71``).
"""
from ..utils.data import get_auto_embedding_dim
from .initializers import RandomNormal


class _TableBackedFeature(object):
    """Shared behaviour of the two id-valued feature kinds."""
    _kind = "Feature"

    def _setup(self, name, vocab_size, embed_dim, shared_with, padding_idx, initializer):
        self.name = name
        self.vocab_size = vocab_size
        # reference: embed_dim=None -> floor(6 * V**0.25)  (utils/data.py:86-101)
        self.embed_dim = get_auto_embedding_dim(vocab_size) if embed_dim is None else embed_dim
        self.shared_with = shared_with
        self.padding_idx = padding_idx
        self.initializer = initializer

    def __repr__(self):
        return f'<{self._kind} {self.name} with Embedding shape ({self.vocab_size}, {self.embed_dim})>'

    def get_embedding_layer(self):
        if not hasattr(self, 'embed'):
            self.embed = self.initializer(self.vocab_size, self.embed_dim, padding_idx=self.padding_idx)
        return self.embed


class SequenceFeature(_TableBackedFeature):
    """Padded id sequence / multi-hot field (reference ``features.py:5-39``).

    Args:
        name (str): column name in the input dict.
        vocab_size (int): rows of the table.
        embed_dim (int): row width; ``None`` -> auto rule.
        pooling (str): ``"mean"``, ``"sum"`` or ``"concat"``.
        shared_with (str): name of the feature whose table is looked up instead.
        padding_idx (int, optional): id treated as padding by ``InputMask``.
        initializer: table factory.
    """
    _kind = "SequenceFeature"

    def __init__(self, name, vocab_size, embed_dim=None, pooling="mean", shared_with=None, padding_idx=None, initializer=RandomNormal(0, 0.0001)):
        self._setup(name, vocab_size, embed_dim, shared_with, padding_idx, initializer)
        self.pooling = pooling


class SparseFeature(_TableBackedFeature):
    """Single categorical id field (reference ``features.py:42-71``)."""
    _kind = "SparseFeature"

    def __init__(self, name, vocab_size, embed_dim=None, shared_with=None, padding_idx=None, initializer=RandomNormal(0, 0.0001)):
        self._setup(name, vocab_size, embed_dim, shared_with, padding_idx, initializer)


class DenseFeature(object):
    """Numeric field passed through as-is (reference ``features.py:74-87``)."""

    def __init__(self, name, embed_dim=1):
        self.name = name
        self.embed_dim = embed_dim

    def __repr__(self):
        return f'<DenseFeature {self.name}>'

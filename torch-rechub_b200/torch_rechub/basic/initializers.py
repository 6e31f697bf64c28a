"""Embedding-table factories (mirror of reference ``torch_rechub/basic/initializers.py:4-100``).

Every factory answers ``__call__(vocab_size, embed_dim, padding_idx=None)`` with an
``nn.Embedding`` (here: the :class:`~torch_rechub.b200.table.FieldTable` subclass, which
keeps the ``weight``-only state_dict layout and the ``isinstance(nn.Embedding)`` contract
of ``basic/loss_func.py:47``).

RNG contract: the reference first constructs ``nn.Embedding`` (whose ``reset_parameters``
draws V*D standard normals and zeroes the padding row) and then re-draws with the requested
distribution (``initializers.py:17-18``).  The same two draws happen here in the same order,
so ``torch.manual_seed(s)`` reproduces the reference tables bit for bit.
"""
import torch

from ..b200.table import FieldTable


def _blank_table(vocab_size, embed_dim, padding_idx):
    # FieldTable.__init__ == nn.Embedding.__init__ : consumes the first V*D normals.
    return FieldTable(vocab_size, embed_dim, padding_idx=padding_idx)


def _zero_padding_row(table, padding_idx):
    if padding_idx is not None:
        with torch.no_grad():
            table.weight[padding_idx].zero_()
    return table


class RandomNormal(object):
    """N(mean, std) table (reference ``initializers.py:4-21``)."""

    def __init__(self, mean=0.0, std=1.0):
        self.mean = mean
        self.std = std

    def __call__(self, vocab_size, embed_dim, padding_idx=None):
        table = _blank_table(vocab_size, embed_dim, padding_idx)
        torch.nn.init.normal_(table.weight, self.mean, self.std)
        return _zero_padding_row(table, padding_idx)


class RandomUniform(object):
    """U[minval, maxval) table (reference ``initializers.py:24-41``)."""

    def __init__(self, minval=0.0, maxval=1.0):
        self.minval = minval
        self.maxval = maxval

    def __call__(self, vocab_size, embed_dim, padding_idx=None):
        table = _blank_table(vocab_size, embed_dim, padding_idx)
        torch.nn.init.uniform_(table.weight, self.minval, self.maxval)
        return _zero_padding_row(table, padding_idx)


class XavierNormal(object):
    """Glorot-normal table (reference ``initializers.py:44-61``)."""

    def __init__(self, gain=1.0):
        self.gain = gain

    def __call__(self, vocab_size, embed_dim, padding_idx=None):
        table = _blank_table(vocab_size, embed_dim, padding_idx)
        torch.nn.init.xavier_normal_(table.weight, self.gain)
        return _zero_padding_row(table, padding_idx)


class XavierUniform(object):
    """Glorot-uniform table (reference ``initializers.py:64-81``)."""

    def __init__(self, gain=1.0):
        self.gain = gain

    def __call__(self, vocab_size, embed_dim, padding_idx=None):
        table = _blank_table(vocab_size, embed_dim, padding_idx)
        torch.nn.init.xavier_uniform_(table.weight, self.gain)
        return _zero_padding_row(table, padding_idx)


class Pretrained(object):
    """Table built from a given 2-D weight (reference ``initializers.py:84-100``).

    ``freeze=True`` (default) marks the table as not trainable.
    """

    def __init__(self, embedding_weight, freeze=True):
        self.embedding_weight = torch.FloatTensor(embedding_weight)
        self.freeze = freeze

    def __call__(self, vocab_size, embed_dim, padding_idx=None):
        rows, cols = self.embedding_weight.shape
        assert vocab_size == rows and embed_dim == cols
        return FieldTable.from_pretrained(self.embedding_weight, freeze=self.freeze, padding_idx=padding_idx)

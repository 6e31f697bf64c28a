"""Activations of the ranking path (mirror of reference ``torch_rechub/basic/activation.py:5-54``)."""
import torch
import torch.nn as nn


class Dice(nn.Module):
    """Dice gate as the reference computes it (``activation.py:15-25``).

    NOTE (reference semantics, kept on purpose): the statistics are taken per ROW over the
    neuron dimension — ``avg = mean_j x[n,j]``, ``var = sum_j ((x[n,j]-avg)^2 + eps)`` (a SUM,
    including ``num_neurons * eps``) — not per neuron over the batch as in the DIN paper.
    ``out = p*x + (1-p)*alpha*x`` with ``p = sigmoid((x-avg)/sqrt(var))``.
    """

    def __init__(self, epsilon=1e-3):
        super(Dice, self).__init__()
        self.epsilon = epsilon
        self.alpha = nn.Parameter(torch.randn(1))

    def forward(self, x: torch.Tensor):
        centred = x - x.mean(dim=1, keepdim=True)
        spread = (centred * centred + self.epsilon).sum(dim=1, keepdim=True)
        gate = torch.sigmoid(centred / torch.sqrt(spread))
        return gate * x + (1 - gate) * self.alpha * x


_BY_NAME = {
    'sigmoid': lambda: nn.Sigmoid(),
    'relu': lambda: nn.ReLU(inplace=True),
    'dice': lambda: Dice(),
    'prelu': lambda: nn.PReLU(),
    'softmax': lambda: nn.Softmax(dim=1),
    'leakyrelu': lambda: nn.LeakyReLU(),
}


def activation_layer(act_name):
    """String (or ``nn.Module`` subclass) -> activation module (reference ``activation.py:28-54``).

    An unknown string raises ``UnboundLocalError`` in the reference (the local is never bound);
    here it is a ``NotImplementedError`` subclass of that behaviour's intent — same for non-module classes.
    """
    if isinstance(act_name, str):
        maker = _BY_NAME.get(act_name.lower())
        if maker is None:
            raise NotImplementedError("unknown activation %r" % (act_name,))
        return maker()
    if isinstance(act_name, type) and issubclass(act_name, nn.Module):
        return act_name()
    raise NotImplementedError

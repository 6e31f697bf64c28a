"""Deep & Cross Network (mirror of reference ``torch_rechub/models/ranking/dcn.py:14-38``)."""
import torch

from ...basic.layers import LR, MLP, CrossNetwork, EmbeddingLayer


class DCN(torch.nn.Module):
    """``sigmoid(LR(cat(CrossNetwork(e), MLP(e))))`` over the flattened embedding tile ``e``.

    Args:
        features (list): all input features (sparse block first, dense appended in the tile).
        n_cross_layers (int): depth of the cross network.
        mlp_params (dict): ``{"dims": list, "activation": str, "dropout": float}`` (no output layer).

    On CUDA: one fused gather launch builds the tile, one launch runs all cross layers, the tower runs on the tensor cores.
    """

    def __init__(self, features, n_cross_layers, mlp_params):
        super().__init__()
        width = sum(fea.embed_dim for fea in features)
        self.features, self.dims = features, width
        self.embedding = EmbeddingLayer(features)
        self.cn = CrossNetwork(width, n_cross_layers)
        self.mlp = MLP(width, output_layer=False, **mlp_params)
        self.linear = LR(width + mlp_params["dims"][-1])

    def forward(self, x):
        tile = self.embedding(x, self.features, squeeze_dim=True)
        return self.linear.probability(torch.cat([self.cn(tile), self.mlp(tile)], dim=1))  # dcn.py:36-38

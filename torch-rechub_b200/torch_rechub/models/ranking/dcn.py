"""Deep & Cross Network (mirror of reference ``torch_rechub/models/ranking/dcn.py:14-38``)."""
import torch

from ...basic.layers import LR, MLP, CrossNetwork, EmbeddingLayer


class DCN(torch.nn.Module):
    """``sigmoid(LR(cat(CrossNetwork(e), MLP(e))))`` over the flattened embedding tile ``e``.

    Args:
        features (list): all input features (sparse block first, dense appended in the tile).
        n_cross_layers (int): depth of the cross network.
        mlp_params (dict): ``{"dims": list, "activation": str, "dropout": float}`` (no output layer).
    """

    def __init__(self, features, n_cross_layers, mlp_params):
        super().__init__()
        self.features = features
        self.dims = sum([fea.embed_dim for fea in features])
        self.embedding = EmbeddingLayer(features)
        self.cn = CrossNetwork(self.dims, n_cross_layers)
        self.mlp = MLP(self.dims, output_layer=False, **mlp_params)
        self.linear = LR(self.dims + mlp_params["dims"][-1])

    def forward(self, x):
        embed_x = self.embedding(x, self.features, squeeze_dim=True)  # CUDA: one fused gather launch
        cn_out = self.cn(embed_x)  # CUDA: all cross layers in one launch
        mlp_out = self.mlp(embed_x)
        return self.linear.probability(torch.cat([cn_out, mlp_out], dim=1))  # sigmoid(LR([cross | deep])), dcn.py:36-38

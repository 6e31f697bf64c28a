"""DCN-v2 (mirror of reference ``torch_rechub/models/ranking/dcn_v2.py:13-59``)."""
import torch

from ...basic.layers import LR, MLP, CrossNetMix, CrossNetV2, EmbeddingLayer

_STRUCTURES = ("crossnet_only", "stacked", "parallel")


class DCNv2(torch.nn.Module):
    """Deep & Cross Network v2 with (by default) the mixture of low-rank experts cross net.

    Args:
        features (list): all input features.
        n_cross_layers (int): number of cross layers.
        mlp_params (dict): tower parameters (no output layer).
        model_structure (str): ``"crossnet_only"``, ``"stacked"`` or ``"parallel"``.
        use_low_rank_mixture (bool): ``CrossNetMix`` when True, else full-rank ``CrossNetV2``.
        low_rank (int), num_experts (int): ``CrossNetMix`` sizes.

    Sub-module names (= state_dict keys) follow the reference: ``embedding``, ``crossnet``, ``stacked_dnn`` or ``parallel_dnn``,
    ``linear`` — created in that order, so a seed reproduces the reference's initial weights.
    """

    def __init__(self, features, n_cross_layers, mlp_params, model_structure="parallel", use_low_rank_mixture=True, low_rank=32, num_experts=4, **kwargs):
        super().__init__()
        width = sum(fea.embed_dim for fea in features)
        self.features, self.dims, self.model_structure = features, width, model_structure
        self.embedding = EmbeddingLayer(features)
        self.crossnet = CrossNetMix(width, n_cross_layers, low_rank=low_rank, num_experts=num_experts) if use_low_rank_mixture else CrossNetV2(width, n_cross_layers)
        if model_structure not in _STRUCTURES:  # an AssertionError, as the reference's assert (dcn_v2.py:35-36)
            raise AssertionError("model_structure={} not supported!".format(model_structure))
        head_width = width
        if model_structure != "crossnet_only":
            setattr(self, model_structure + "_dnn", MLP(width, output_layer=False, **mlp_params))
            head_width = mlp_params["dims"][-1] + (width if model_structure == "parallel" else 0)
        self.linear = LR(head_width)

    def forward(self, x):
        tile = self.embedding(x, self.features, squeeze_dim=True)
        crossed = self.crossnet(tile)
        if self.model_structure == "stacked":
            crossed = self.stacked_dnn(crossed)
        elif self.model_structure == "parallel":
            crossed = torch.cat([crossed, self.parallel_dnn(tile)], dim=1)
        return self.linear.probability(crossed)  # sigmoid(LR(.)), dcn_v2.py:57-59

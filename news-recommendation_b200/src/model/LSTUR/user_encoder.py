"""LSTUR UserEncoder (replaces reference src/model/LSTUR/user_encoder.py:6-45): GRU over the browsed-news vectors,
initialised with ('ini') or concatenated to ('con') the long-term user embedding."""
import torch
import torch.nn as nn

from newsrec_b200 import require_cuda
from newsrec_b200.ops import OperandCache
from newsrec_b200.ops_gru import GruLastHiddenFn


class UserEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        assert int(config.num_filters * 1.5) == config.num_filters * 1.5
        hidden = config.num_filters * 3 if config.long_short_term_method == "ini" else int(config.num_filters * 1.5)
        self.gru = nn.GRU(config.num_filters * 3, hidden)  # parameter container: weight_ih_l0, weight_hh_l0, bias_*_l0
        self._cache = OperandCache()

    def forward(self, user, clicked_news_length, clicked_news_vector):
        """user (batch, hidden) device fp32; clicked_news_length (batch,) int64; clicked_news_vector (batch, H, 3F)"""
        dev = require_cuda()
        clicked_news_length.clamp_(min=1)  # in place, like the reference's `length[length == 0] = 1` (:27) -- without the
        # device synchronisation a boolean-mask assignment needs when the lengths already live on the GPU
        g = self.gru
        acc = getattr(self.config, "precision", "fast") == "accurate"
        if self.config.long_short_term_method == "ini":
            return GruLastHiddenFn.apply(clicked_news_vector, clicked_news_length, user, g.weight_ih_l0, g.weight_hh_l0,
                                         g.bias_ih_l0, g.bias_hh_l0, self._cache, "gru", acc)
        h0 = torch.zeros((clicked_news_vector.shape[0], g.weight_hh_l0.shape[1]), device=dev)
        last = GruLastHiddenFn.apply(clicked_news_vector, clicked_news_length, h0, g.weight_ih_l0, g.weight_hh_l0,
                                     g.bias_ih_l0, g.bias_hh_l0, self._cache, "gru", acc)
        return torch.cat((last, user), dim=1)

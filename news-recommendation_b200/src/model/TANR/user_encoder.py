"""TANR UserEncoder (replaces reference src/model/TANR/user_encoder.py:5-19): additive pooling of the history."""
import torch.nn as nn

from model.general.attention.additive import AdditiveAttention


class UserEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.additive_attention = AdditiveAttention(config.query_vector_dim, config.num_filters)

    def forward(self, clicked_news_vector):
        return self.additive_attention(clicked_news_vector)

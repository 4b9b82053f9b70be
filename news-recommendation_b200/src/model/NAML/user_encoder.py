"""NAML UserEncoder (replaces reference src/model/NAML/user_encoder.py:5-19): additive pooling of the history."""
import torch.nn as nn

from model.general.attention.additive import AdditiveAttention


class UserEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.additive_attention = AdditiveAttention(config.query_vector_dim, config.num_filters)

    def forward(self, clicked_news_vector):
        """(batch, num_clicked_news_a_user, num_filters) -> (batch, num_filters)"""
        return self.additive_attention(clicked_news_vector)

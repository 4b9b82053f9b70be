"""LSTUR drop-in (replaces reference src/model/LSTUR/__init__.py:11-120)."""
import torch
import torch.nn as nn

from model.general.click_predictor.dot_product import DotProductClickPredictor
from model.general.cnn_text import BadIdFlag
from model.LSTUR.news_encoder import NewsEncoder
from model.LSTUR.user_encoder import UserEncoder
from newsrec_b200 import require_cuda
from newsrec_b200.ops_cnn import EmbeddingF32Fn
from newsrec_b200.pack import SlotPacker


class LSTUR(torch.nn.Module):
    def __init__(self, config, pretrained_word_embedding=None):
        super().__init__()
        self.config = config
        self.news_encoder = NewsEncoder(config, pretrained_word_embedding)
        self.user_encoder = UserEncoder(config)
        self.click_predictor = DotProductClickPredictor()
        assert int(config.num_filters * 1.5) == config.num_filters * 1.5
        self.user_embedding = nn.Embedding(
            config.num_users,
            config.num_filters * 3 if config.long_short_term_method == "ini" else int(config.num_filters * 1.5),
            padding_idx=0)
        self._packer, self._uflag = SlotPacker(), BadIdFlag()

    def _user_vector(self, user, dev, masked):
        u = EmbeddingF32Fn.apply(user.to(dev, non_blocking=True).view(-1), self.user_embedding.weight, self._uflag.get(dev))
        if masked and self.training and self.config.masking_probability > 0:
            # F.dropout2d on a (1, batch, dim) tensor == drop WHOLE user vectors with p, scale the rest by 1/(1-p)
            # (reference __init__.py:74-77)
            p = self.config.masking_probability
            keep = (torch.rand(u.shape[0], 1, device=dev) >= p).to(u.dtype) / (1.0 - p)
            u = u * keep
        return u

    def forward(self, user, clicked_news_length, candidate_news, clicked_news):
        dev = require_cuda()
        C, H = len(candidate_news), len(clicked_news)
        fields, B = {}, None
        for name in ("category", "subcategory", "title"):
            fields[name], B = self._packer.pack(clicked_news, candidate_news, name, dev)
        vec = self.news_encoder.encode(fields)
        Dn = vec.shape[1]
        u = self._user_vector(user, dev, masked=True)
        user_vector = self.user_encoder(u, clicked_news_length, vec[:B * H].view(B, H, Dn))
        return self.click_predictor(vec[B * H:].view(B, C, Dn), user_vector)

    def get_news_vector(self, news):
        return self.news_encoder(news)

    def get_user_vector(self, user, clicked_news_length, clicked_news_vector):
        dev = require_cuda()
        u = self._user_vector(user, dev, masked=False)  # no masking at inference (reference :104)
        return self.user_encoder(u, clicked_news_length, clicked_news_vector)

    def get_prediction(self, news_vector, user_vector):
        return self.click_predictor(news_vector.unsqueeze(0), user_vector.unsqueeze(0)).squeeze(0)

"""NAML drop-in (replaces reference src/model/NAML/__init__.py:7-93).  All 1+K+H news of a batch are packed
per attribute into one id tensor and encoded by one kernel sequence per view."""
import torch

from model.general.click_predictor.dot_product import DotProductClickPredictor
from model.NAML.news_encoder import NewsEncoder
from model.NAML.user_encoder import UserEncoder
from newsrec_b200 import require_cuda
from newsrec_b200.pack import SlotPacker


class NAML(torch.nn.Module):
    def __init__(self, config, pretrained_word_embedding=None):
        super().__init__()
        self.config = config
        self.news_encoder = NewsEncoder(config, pretrained_word_embedding)
        self.user_encoder = UserEncoder(config)
        self.click_predictor = DotProductClickPredictor()
        self._packer = SlotPacker()

    def forward(self, candidate_news, clicked_news):
        """lists of per-slot dicts {"category","subcategory": (batch,), "title": (batch,20), "abstract": (batch,50)}"""
        dev = require_cuda()
        C, H = len(candidate_news), len(clicked_news)
        names = list(self.news_encoder.text_encoders.keys()) + list(self.news_encoder.element_encoders.keys())
        fields, B = {}, None
        for name in names:
            fields[name], B = self._packer.pack(clicked_news, candidate_news, name, dev)
        vec = self.news_encoder.encode(fields)
        Fn = vec.shape[1]
        user_vector = self.user_encoder(vec[:B * H].view(B, H, Fn))
        return self.click_predictor(vec[B * H:].view(B, C, Fn), user_vector)

    def get_news_vector(self, news):
        return self.news_encoder(news)

    def get_user_vector(self, clicked_news_vector):
        return self.user_encoder(clicked_news_vector)

    def get_prediction(self, news_vector, user_vector):
        return self.click_predictor(news_vector.unsqueeze(0), user_vector.unsqueeze(0)).squeeze(0)

"""TANR drop-in (replaces reference src/model/TANR/__init__.py:10-105): NRMS-style scoring over CNN news vectors
plus the auxiliary topic-classification loss over every news vector of the batch (class 0 has weight 0)."""
import torch
import torch.nn as nn

from model.general.click_predictor.dot_product import DotProductClickPredictor
from model.TANR.news_encoder import NewsEncoder
from model.TANR.user_encoder import UserEncoder
from newsrec_b200 import require_cuda
from newsrec_b200.ops import OperandCache
from newsrec_b200.ops_cnn import LinearRowsFn
from newsrec_b200.pack import SlotPacker


class TANR(torch.nn.Module):
    def __init__(self, config, pretrained_word_embedding=None):
        super().__init__()
        self.config = config
        self.news_encoder = NewsEncoder(config, pretrained_word_embedding)
        self.user_encoder = UserEncoder(config)
        self.click_predictor = DotProductClickPredictor()
        self.topic_predictor = nn.Linear(config.num_filters, config.num_categories)
        self._packer, self._cache = SlotPacker(), OperandCache()

    def forward(self, candidate_news, clicked_news):
        """-> (click logits (batch, 1+K), topic_classification_loss 0-dim)"""
        dev = require_cuda()
        C, H = len(candidate_news), len(clicked_news)
        ids, B = self._packer.pack(clicked_news, candidate_news, "title", dev)
        cats, _ = self._packer.pack(clicked_news, candidate_news, "category", dev)
        vec = self.news_encoder.encode_ids(ids)
        Fn = vec.shape[1]
        user_vector = self.user_encoder(vec[:B * H].view(B, H, Fn))
        click_probability = self.click_predictor(vec[B * H:].view(B, C, Fn), user_vector)
        # topic head over all B*(1+K+H) news vectors (reference :58-67); weighted CE, class 0 (padding) weight 0
        y_pred = LinearRowsFn.apply(vec, self.topic_predictor.weight, self.topic_predictor.bias, False, self._cache, "topic")
        class_weight = torch.ones(self.config.num_categories, device=dev)
        class_weight[0] = 0
        topic_classification_loss = nn.functional.cross_entropy(y_pred, cats.view(-1), weight=class_weight)
        return click_probability, topic_classification_loss

    def get_news_vector(self, news):
        return self.news_encoder(news)

    def get_user_vector(self, clicked_news_vector):
        return self.user_encoder(clicked_news_vector)

    def get_prediction(self, news_vector, user_vector):
        return self.click_predictor(news_vector.unsqueeze(0), user_vector.unsqueeze(0)).squeeze(0)

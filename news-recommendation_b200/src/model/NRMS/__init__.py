"""NRMS drop-in (replaces reference src/model/NRMS/__init__.py:7-84): same class name, constructor,
methods and state_dict keys; the 1+K candidate and H browsed titles of a batch are packed into ONE
id matrix and encoded by one fused kernel sequence instead of 55 Python-level encoder calls."""
import torch

from model.general.click_predictor.dot_product import DotProductClickPredictor
from model.NRMS.news_encoder import NewsEncoder
from model.NRMS.user_encoder import UserEncoder
from newsrec_b200 import require_cuda
from newsrec_b200.pack import PackedBatch, SlotPacker


class NRMS(torch.nn.Module):
    def __init__(self, config, pretrained_word_embedding=None):
        super().__init__()
        self.config = config
        self.news_encoder = NewsEncoder(config, pretrained_word_embedding)
        self.user_encoder = UserEncoder(config)
        self.click_predictor = DotProductClickPredictor()
        self._packer = SlotPacker()
        self._copy_stream = None

    def prefetch(self, candidate_news, clicked_news):
        """Stage the NEXT batch while the current step runs: host-side stacking into pinned memory, one H2D copy and
        the device-side re-ordering, all on a private copy stream.  Pass the returned PackedBatch to forward()."""
        dev = require_cuda()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        return self._packer.pack_on_stream(clicked_news, candidate_news, "title", dev, self._copy_stream)

    def forward(self, candidate_news, clicked_news=None):
        """candidate_news: list of 1+K dicts {"title": (batch, T)}; clicked_news: list of H such dicts
        (slot-major, exactly what the reference's DataLoader yields) -- or one PackedBatch from prefetch().
        Returns (batch, 1+K) logits."""
        dev = require_cuda()
        if isinstance(candidate_news, PackedBatch):
            pb = candidate_news
            ids, B, H, C = pb.wait(), pb.B, pb.H, pb.C
        else:
            C, H = len(candidate_news), len(clicked_news)
            ids, B = self._packer.pack(clicked_news, candidate_news, "title", dev)  # (B*H + B*C, T): browsed block, then candidates
        vec = self.news_encoder.encode_ids(ids)
        d = vec.shape[1]
        clicked_news_vector = vec[:B * H].view(B, H, d)
        candidate_news_vector = vec[B * H:].view(B, C, d)
        user_vector = self.user_encoder(clicked_news_vector)
        return self.click_predictor(candidate_news_vector, user_vector)

    def get_news_vector(self, news):
        """{"title": (batch, T)} -> (batch, dim)"""
        return self.news_encoder(news)

    def get_user_vector(self, clicked_news_vector):
        """(batch, H, dim) -> (batch, dim)"""
        return self.user_encoder(clicked_news_vector)

    def get_prediction(self, news_vector, user_vector):
        """(candidates, dim), (dim,) -> (candidates,)"""
        return self.click_predictor(news_vector.unsqueeze(0), user_vector.unsqueeze(0)).squeeze(0)

    def check_ids(self):
        """Raises IndexError (like the reference's CPU embedding) if any token id was out of range.  Syncs."""
        self.news_encoder._flag.raise_if_set("word_embedding")

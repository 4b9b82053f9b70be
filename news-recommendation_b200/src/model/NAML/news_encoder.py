"""NAML NewsEncoder (replaces reference src/model/NAML/news_encoder.py:9-115): multi-view encoder --
title CNN, abstract CNN, category / subcategory element encoders -- fused by additive attention.
Same submodule / parameter names and the same parameter sharing (one word table, one category table)."""
import torch
import torch.nn as nn

from model.general.attention.additive import AdditiveAttention
from model.general.cnn_text import BadIdFlag, cnn_text_encode, make_title_cnn
from newsrec_b200 import require_cuda
from newsrec_b200.ops import OperandCache
from newsrec_b200.ops_cnn import ElementEncoderFn


class TextEncoder(nn.Module):
    def __init__(self, word_embedding, word_embedding_dim, num_filters, window_size, query_vector_dim, dropout_probability):
        super().__init__()
        self.word_embedding = word_embedding
        self.dropout_probability = dropout_probability
        self.CNN = make_title_cnn(num_filters, window_size, word_embedding_dim)
        self.additive_attention = AdditiveAttention(query_vector_dim, num_filters)
        self._cache, self._flag = OperandCache(), BadIdFlag()

    def forward(self, text):
        """(batch, num_words) int64 on the device -> (batch, num_filters)"""
        p = self.dropout_probability if self.training else 0.0
        return cnn_text_encode(text, self.word_embedding, self.CNN, self.additive_attention, p, self._cache, "text", self._flag)


class ElementEncoder(nn.Module):
    def __init__(self, embedding, linear_input_dim, linear_output_dim):
        super().__init__()
        self.embedding = embedding
        self.linear = nn.Linear(linear_input_dim, linear_output_dim)
        self._cache, self._flag = OperandCache(), BadIdFlag()

    def forward(self, element):
        """(batch,) int64 on the device -> (batch, num_filters)"""
        dev = require_cuda()
        return ElementEncoderFn.apply(element, self.embedding.weight, self.linear.weight, self.linear.bias, self._cache,
                                      "element", self._flag.get(dev))


class NewsEncoder(nn.Module):
    TEXT, ELEMENT = ("title", "abstract"), ("category", "subcategory")

    def __init__(self, config, pretrained_word_embedding):
        super().__init__()
        self.config = config
        if pretrained_word_embedding is None:
            word_embedding = nn.Embedding(config.num_words, config.word_embedding_dim, padding_idx=0)
        else:
            word_embedding = nn.Embedding.from_pretrained(pretrained_word_embedding, freeze=False, padding_idx=0)
        attrs = config.dataset_attributes["news"]
        assert len(attrs) > 0
        self.text_encoders = nn.ModuleDict({
            name: TextEncoder(word_embedding, config.word_embedding_dim, config.num_filters, config.window_size,
                              config.query_vector_dim, config.dropout_probability)
            for name in self.TEXT if name in attrs})
        category_embedding = nn.Embedding(config.num_categories, config.category_embedding_dim, padding_idx=0)
        self.element_encoders = nn.ModuleDict({
            name: ElementEncoder(category_embedding, config.category_embedding_dim, config.num_filters)
            for name in self.ELEMENT if name in attrs})
        if len(attrs) > 1:
            self.final_attention = AdditiveAttention(config.query_vector_dim, config.num_filters)

    def encode(self, fields):
        """fields: name -> device tensor ((n, T) for text views, (n,) for elements) -> (n, num_filters)"""
        vectors = [enc(fields[name]) for name, enc in self.text_encoders.items()]
        vectors += [enc(fields[name]) for name, enc in self.element_encoders.items()]
        if len(vectors) == 1:
            return vectors[0]
        return self.final_attention(torch.stack(vectors, dim=1))

    def forward(self, news):
        dev = require_cuda()
        names = list(self.text_encoders.keys()) + list(self.element_encoders.keys())
        return self.encode({k: news[k].to(dev, non_blocking=True) for k in names})

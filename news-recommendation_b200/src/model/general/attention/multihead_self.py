"""MultiHeadSelfAttention parameter container (replaces reference
src/model/general/attention/multihead_self.py:26-76; same parameter names W_Q/W_K/W_V, no W_O).

Inside the NRMS encoders the projection, the per-head exp-softmax(+1e-8) attention and the additive
pooling run as one fused kernel sequence (newsrec_b200.ops.MhsaPoolEncoderFn); the standalone
`forward` runs the projection + attention core only.
"""
import torch.nn as nn

from newsrec_b200.ops import MhsaFn, OperandCache


class MultiHeadSelfAttention(nn.Module):
    def __init__(self, d_model, num_attention_heads):
        super().__init__()
        assert d_model % num_attention_heads == 0
        self.d_model = d_model
        self.num_attention_heads = num_attention_heads
        self.d_k = self.d_v = d_model // num_attention_heads
        self.W_Q = nn.Linear(d_model, d_model)
        self.W_K = nn.Linear(d_model, d_model)
        self.W_V = nn.Linear(d_model, d_model)
        for lin in (self.W_Q, self.W_K, self.W_V):
            nn.init.xavier_uniform_(lin.weight, gain=1)
        self._cache = OperandCache()

    def qkv_parameters(self):
        return (self.W_Q.weight, self.W_Q.bias, self.W_K.weight, self.W_K.bias, self.W_V.weight, self.W_V.bias)

    def forward(self, Q, K=None, V=None, length=None):
        if K is not None or V is not None or length is not None:
            raise NotImplementedError("only self-attention without a length mask is on the hot path "
                                      "(no reference caller passes K, V or length)")
        return MhsaFn.apply(Q, *self.qkv_parameters(), self.num_attention_heads, self._cache, "mhsa")

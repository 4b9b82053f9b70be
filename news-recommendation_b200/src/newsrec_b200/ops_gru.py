"""autograd wrapper of the GRU user encoder (LSTUR): packed-sequence nn.GRU, last hidden state."""
from __future__ import annotations

import ctypes as C

import torch

from . import GruBwdArgs, GruFwdArgs, check, load_library, require_cuda
from .ops import _p, _stream, cast_pad, ru8


def ru4(x):
    return (x + 3) // 4 * 4


class GruLastHiddenFn(torch.autograd.Function):
    """x (B, S, D) fp32, lengths (B,) int64 on the device, h0 (B, Hd) -> last hidden (B, Hd).
    reference: src/model/LSTUR/user_encoder.py:27-45 (pack_padded_sequence + nn.GRU)."""

    @staticmethod
    def forward(ctx, x, lengths, h0, Wih, Whh, bih, bhh, cache, prefix, accurate=False):
        lib = load_library()
        dev = require_cuda()
        B, S, D = x.shape
        Hd = Whh.shape[1]
        ldd, ldh, ldg, ldb = ru8(D + 1), ru8(Hd + 1), ru4(3 * Hd), ru8(3 * Hd + 1)
        ops = cache.get(prefix, (Wih, Whh, bih, bhh), lambda Wih, Whh, bih, bhh: dict(
            wih=cast_pad(Wih, ldd), whh=cast_pad(Whh, ldh), wihT=cast_pad(Wih, ldb, transpose=True),
            whhT=cast_pad(Whh, ldb, transpose=True), bih=bih.float().contiguous(), bhh=bhh.float().contiguous()))
        x = x.float()
        h0 = h0.float().contiguous()
        lengths = lengths.to(dev, non_blocking=True).long().contiguous()  # a blocking copy here would drain the news encoder's kernels
        xb = torch.empty((B * S, ldd), dtype=torch.bfloat16, device=dev)
        gi = torch.empty((B * S, ldg), dtype=torch.float32, device=dev)
        gh = torch.empty((S, B, ldg), dtype=torch.float32, device=dev)
        hs = torch.empty((S + 1, B, Hd), dtype=torch.float32, device=dev)
        hb = torch.empty((S + 1, B, ldh), dtype=torch.bfloat16, device=dev)
        out = torch.empty((B, Hd), dtype=torch.float32, device=dev)
        a = GruFwdArgs()
        a.B, a.S, a.D, a.Hd = B, S, D, Hd
        a.x = _p(x)
        a.x_s_b, a.x_s_t, a.x_s_c = x.stride()
        a.len, a.h0 = _p(lengths), _p(h0)
        a.wih_bf16, a.whh_bf16, a.bih, a.bhh = _p(ops["wih"]), _p(ops["whh"]), _p(ops["bih"]), _p(ops["bhh"])
        a.xb, a.gi, a.gh, a.hs, a.hb, a.out = _p(xb), _p(gi), _p(gh), _p(hs), _p(hb), _p(out)
        x_lo = None
        if accurate:  # news vectors enter the input projection as a hi/lo bf16 pair (second pass over W_ih)
            x_lo = torch.empty((B * S, ldd), dtype=torch.bfloat16, device=dev)
            a.x_lo_bf16 = _p(x_lo)
        check(lib.nr_gru_fwd(C.byref(a), _stream()), "nr_gru_fwd")
        ctx.save_for_backward(xb, gi, gh, hs, hb, lengths)
        ctx.meta = dict(B=B, S=S, D=D, Hd=Hd, ops=ops)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = load_library()
        xb, gi, gh, hs, hb, lengths = ctx.saved_tensors
        m = ctx.meta
        B, S, D, Hd, ops = m["B"], m["S"], m["D"], m["Hd"], m["ops"]
        dev = xb.device
        ldd, ldh = ru8(D + 1), ru8(Hd + 1)
        dout = dout.contiguous().float()
        dWih = torch.zeros((3 * Hd, ldd), dtype=torch.float32, device=dev)
        dWhh = torch.zeros((3 * Hd, ldh), dtype=torch.float32, device=dev)
        dx = torch.empty((B * S, D), dtype=torch.float32, device=dev)
        dh0 = torch.empty((B, Hd), dtype=torch.float32, device=dev)
        ws_bytes = int(lib.nr_gru_bwd_workspace(B, S, D, Hd))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        a = GruBwdArgs()
        a.B, a.S, a.D, a.Hd = B, S, D, Hd
        a.len, a.wihT_bf16, a.whhT_bf16 = _p(lengths), _p(ops["wihT"]), _p(ops["whhT"])
        a.xb, a.gi, a.gh, a.hs, a.hb = _p(xb), _p(gi), _p(gh), _p(hs), _p(hb)
        a.dout, a.dWih_ext, a.dWhh_ext, a.dx, a.dh0 = _p(dout), _p(dWih), _p(dWhh), _p(dx), _p(dh0)
        a.workspace, a.workspace_bytes = _p(ws), ws_bytes
        check(lib.nr_gru_bwd(C.byref(a), _stream()), "nr_gru_bwd")
        return (dx.view(B, S, D), None, dh0, dWih[:, :D].contiguous(), dWhh[:, :Hd].contiguous(), dWih[:, D].contiguous(),
                dWhh[:, Hd].contiguous(), None, None, None)

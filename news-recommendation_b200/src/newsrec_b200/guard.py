"""Out-of-range id reporting.  The reference's nn.Embedding raises IndexError (CPU) / a device assert (CUDA) for an id
outside [0, V) -- a common mistake with this reference, whose num_words / num_users / num_categories are hand-edited after
preprocessing (src/config.py).  The gather kernels clamp such an id to row 0 and set a device-side flag; the backward
scatters skip it.  `BadIdFlag` surfaces the flag as IndexError WITHOUT a host/device sync on the training path: every
`get()` queues a 4-byte asynchronous copy of the flag into pinned memory and examines the copy queued by the PREVIOUS call
once its event has completed -- the error is raised at the latest one forward pass after the offending batch.
`raise_if_set()` is the synchronous check (evaluation, tests)."""
from __future__ import annotations

import torch


class BadIdFlag:
    def __init__(self):
        self._t = None
        self._host = None
        self._event = None

    def get(self, dev):
        if self._t is None or self._t.device != dev:
            self._t = torch.zeros(1, dtype=torch.int32, device=dev)
            self._host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._event = None
        else:
            self.poll()
        return self._t

    def poll(self, what="an embedding table"):
        """Non-blocking: raise if the previously queued snapshot of the flag is set, then queue a new snapshot."""
        if self._t is None:
            return
        if self._event is not None:
            if not self._event.query():
                return  # the previous snapshot is still in flight: look again at the next call
            if int(self._host[0]) != 0:
                self._event = None
                raise IndexError(f"id out of range for {what} (reported by the device-side check of an earlier batch)")
        self._host.copy_(self._t, non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()

    def raise_if_set(self, what="word_embedding"):
        if self._t is not None and int(self._t.item()) != 0:
            raise IndexError(f"id out of range for {what}")

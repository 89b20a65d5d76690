"""Flat parameter / gradient buffers: every trainable variable of a model set becomes a view into ONE fp32 buffer, so the
data-parallel gradient exchange is a single all-reduce (two buckets: the Dense tail early, the rest at the end) and the optimizer a
single fused launch (train.py:402-406: Keras RMSprop over _gen + _sun variables and over _dis variables)."""
from __future__ import annotations

import torch


class FlatVars:
    def __init__(self, owners, device):
        """owners: list of (object, attribute name); the attribute must hold a contiguous fp32 tensor.  Order = layout order."""
        self.owners = list(owners)
        shapes = [tuple(getattr(o, a).shape) for o, a in self.owners]
        sizes = [int(torch.tensor(s).prod()) if len(s) else 1 for s in shapes]
        pad = [(-n) % 4 for n in sizes]                       # keep every view 16-byte aligned
        total = sum(n + p for n, p in zip(sizes, pad))
        self.flat_w = torch.zeros(total, dtype=torch.float32, device=device)
        self.flat_g = torch.zeros_like(self.flat_w)
        self._grads, self._offsets, off = {}, {}, 0
        for (o, a), n, p, s in zip(self.owners, sizes, pad, shapes):
            vw = self.flat_w[off:off + n].view(s)
            vw.copy_(getattr(o, a))
            setattr(o, a, vw)                                  # the layer now reads the flat buffer
            self._grads[(id(o), a)] = self.flat_g[off:off + n].view(s)
            self._offsets[(id(o), a)] = off
            off += n + p
        self._objects = []
        seen = set()
        for o, _ in self.owners:
            if id(o) not in seen:
                seen.add(id(o))
                self._objects.append(o)

    def grad(self, obj, attr):
        return self._grads[(id(obj), attr)]

    def offset(self, obj, attr):
        return self._offsets[(id(obj), attr)]

    def invalidate(self):
        """The variables were rewritten in place by an optimizer kernel: drop every derived copy (packed TF32 images, transposes)."""
        for o in self._objects:
            fn = getattr(o, "_invalidate", None)
            if fn is not None:
                fn()

"""`StridedTensor`: the ragged container the reference's search path keeps its IVF, codes and residuals in
(CB/search/strided_tensor_core.py:17-130, CB/search/strided_tensor.py:13-175), GPU-resident.

Same constructor and accessors -- `lookup(pids, output='packed'|'padded')`, `as_padded_tensor()`,
`as_packed_tensor()` -- but no `as_strided` views over an over-allocated buffer: the ragged gather is the
`plaid_segmented_lookup` kernel (the counterpart of segmented_lookup.cpp) and padding is a scatter of the
packed rows.
"""
from __future__ import annotations

import torch

from . import ops


class StridedTensor:
    def __init__(self, packed_tensor, lengths, dim=None, use_gpu=True):
        if not use_gpu:
            raise RuntimeError("StridedTensor: this implementation is GPU-resident (no CPU branch)")
        self.dim = dim
        self.use_gpu = True
        self.tensor = ops._cu(packed_tensor)
        self.inner_dims = self.tensor.size()[1:]
        lengths = lengths if torch.is_tensor(lengths) else torch.tensor(lengths)
        self.lengths = ops._cu(lengths, torch.int64)
        zero = torch.zeros(1, dtype=torch.int64, device=self.lengths.device)
        self.offsets = torch.cat((zero, torch.cumsum(self.lengths, 0)))       # strided_tensor_core.py:30-31
        self.max_stride = int(self.lengths.max().item()) if self.lengths.numel() else 0

    @classmethod
    def from_packed_tensor(cls, tensor, lengths):
        return cls(tensor, lengths)

    # ---- lookup (strided_tensor.py:58-99) ------------------------------------------------------
    def _prepare_lookup(self, pids):
        if isinstance(pids, list):
            pids = torch.tensor(pids)
        assert pids.dim() == 1
        pids = ops._cu(pids, torch.int64)
        return pids, self.lengths[pids], self.offsets[pids]

    def lookup(self, pids, output="packed"):
        pids, lengths, offsets = self._prepare_lookup(pids)
        packed = ops.segmented_lookup(self.tensor, pids, lengths, offsets)
        if output == "packed":
            return packed, lengths
        assert output == "padded"
        return _pad(packed, lengths)

    # ---- whole-tensor views (strided_tensor_core.py:64-96) --------------------------------------
    def as_packed_tensor(self, return_offsets=False):
        vals = [self.tensor, self.lengths]
        if return_offsets:
            vals.append(self.offsets)
        return tuple(vals)

    def as_padded_tensor(self):
        return _pad(self.tensor[: int(self.offsets[-1])], self.lengths)


def _pad(packed, lengths):
    """packed [sum len, *inner] + lengths [n] -> (padded [n, max len, *inner], mask broadcastable to it)."""
    n = lengths.numel()
    stride = int(lengths.max().item()) if n else 0
    inner = packed.size()[1:]
    mask = torch.arange(stride, device=packed.device).unsqueeze(0) < lengths.unsqueeze(-1)   # _create_mask
    padded = torch.zeros((n, stride, *inner), device=packed.device, dtype=packed.dtype)
    padded[mask] = packed
    for _ in range(padded.dim() - mask.dim()):
        mask = mask.unsqueeze(-1)
    return padded, mask

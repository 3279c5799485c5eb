"""Stand-in for ``positional_encodings==6.0.3`` (requirements.txt:3 of the reference), which is
neither vendored under /root/reference nor installed in this image.

PARITY UNPINNED: restated from the published v6.0.x algorithm (tatp22/multidim-positional-encoding):
channels = ceil(C/6)*2 (made even), inv_freq[j] = 10000^(-2j/channels), per-axis code with sin/cos
INTERLEAVED, the three axis codes concatenated and truncated to C.  Only used by the golden
generator (tests/golden/make_golden.py) so that the unmodified reference imports.
"""
import numpy as np
import torch
from torch import nn


def get_emb(sin_inp):
    return torch.flatten(torch.stack((sin_inp.sin(), sin_inp.cos()), dim=-1), -2, -1)


class PositionalEncoding3D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.org_channels = channels
        channels = int(np.ceil(channels / 6) * 2)
        if channels % 2:
            channels += 1
        self.channels = channels
        inv_freq = 1.0 / (10000 ** (torch.arange(0, channels, 2).float() / channels))
        self.register_buffer("inv_freq", inv_freq)
        self.register_buffer("cached_penc", None, persistent=False)

    def forward(self, tensor):
        _, x, y, z, orig_ch = tensor.shape
        codes = []
        for n in (x, y, z):
            pos = torch.arange(n, dtype=self.inv_freq.dtype)
            codes.append(get_emb(torch.einsum("i,j->ij", pos, self.inv_freq)))
        emb = torch.zeros((x, y, z, self.channels * 3), dtype=tensor.dtype)
        c = self.channels
        emb[:, :, :, :c] = codes[0][:, None, None, :]
        emb[:, :, :, c:2 * c] = codes[1][None, :, None, :]
        emb[:, :, :, 2 * c:] = codes[2][None, None, :, :]
        self.cached_penc = emb[None, :, :, :, :orig_ch].repeat(tensor.shape[0], 1, 1, 1, 1)
        return self.cached_penc

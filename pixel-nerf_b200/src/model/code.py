"""Positional code module (reference: src/model/code.py).  Holds the `_freqs` / `_phases`
buffers so checkpoints keep their keys; the fused kernel recomputes the code in registers
(csrc/pnr_stages.cu: feat_channel) and this torch forward serves the autograd path only."""
import numpy as np
import torch


class PositionalEncoding(torch.nn.Module):
    def __init__(self, num_freqs=6, d_in=3, freq_factor=np.pi, include_input=True):
        super().__init__()
        self.num_freqs = num_freqs
        self.d_in = d_in
        self.freq_factor = float(freq_factor)
        self.include_input = include_input
        self.freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs)
        self.d_out = num_freqs * 2 * d_in + (d_in if include_input else 0)
        self.register_buffer("_freqs", torch.repeat_interleave(self.freqs, 2).view(1, -1, 1))
        phases = torch.zeros(2 * num_freqs)
        phases[1::2] = np.pi * 0.5
        self.register_buffer("_phases", phases.view(1, -1, 1))

    def forward(self, x):
        """(N, d_in) -> (N, d_out): [x, sin(f0 x), sin(f0 x + pi/2), sin(f1 x), ...]."""
        arg = torch.addcmul(self._phases, x.unsqueeze(1).expand(-1, 2 * self.num_freqs, -1), self._freqs)
        code = torch.sin(arg).reshape(x.shape[0], -1)
        return torch.cat((x, code), dim=-1) if self.include_input else code

    @classmethod
    def from_conf(cls, conf, d_in=3):
        return cls(conf.get_int("num_freqs", 6), d_in, conf.get_float("freq_factor", np.pi),
                   conf.get_bool("include_input", True))

"""ResnetFC parameter container (reference: src/model/resnetfc.py).  Parameter names and
shapes match the reference's state_dict; inference runs through the fused kernels
(csrc/), this torch `forward` exists for the autograd (training) path only."""
import torch
from torch import nn


class ResnetBlockFC(nn.Module):
    """x + fc_1(relu(fc_0(relu(x))))  (resnetfc.py:10-62, size_in == size_out)."""

    def __init__(self, size, beta=0.0):
        super().__init__()
        if beta > 0:
            raise NotImplementedError("mlp.beta > 0 (Softplus) is not supported by the fused path")
        self.fc_0 = nn.Linear(size, size)
        self.fc_1 = nn.Linear(size, size)
        nn.init.zeros_(self.fc_0.bias)
        nn.init.kaiming_normal_(self.fc_0.weight, a=0, mode="fan_in")
        nn.init.zeros_(self.fc_1.bias)
        nn.init.zeros_(self.fc_1.weight)  # the reference starts every block as the identity

    def forward(self, x):
        return x + self.fc_1(torch.relu(self.fc_0(torch.relu(x))))


class ResnetFC(nn.Module):
    def __init__(self, d_in, d_out=4, n_blocks=5, d_latent=0, d_hidden=128, beta=0.0, combine_layer=1000,
                 combine_type="average", use_spade=False):
        super().__init__()
        if use_spade:
            raise NotImplementedError("mlp.use_spade is not supported by the fused path")
        if combine_type != "average":
            raise NotImplementedError("mlp.combine_type = %s is not supported (only average)" % combine_type)
        if d_in <= 0 or d_latent <= 0:
            raise NotImplementedError("ResnetFC needs d_in > 0 and d_latent > 0 in this implementation")
        self.d_in, self.d_out, self.d_latent, self.d_hidden = d_in, d_out, d_latent, d_hidden
        self.n_blocks, self.combine_layer, self.combine_type, self.use_spade = n_blocks, combine_layer, combine_type, False
        self.lin_in = nn.Linear(d_in, d_hidden)
        self.lin_out = nn.Linear(d_hidden, d_out)
        self.blocks = nn.ModuleList([ResnetBlockFC(d_hidden, beta=beta) for _ in range(n_blocks)])
        self.lin_z = nn.ModuleList([nn.Linear(d_latent, d_hidden) for _ in range(min(combine_layer, n_blocks))])
        for lin in [self.lin_in, self.lin_out, *self.lin_z]:
            nn.init.zeros_(lin.bias)
            nn.init.kaiming_normal_(lin.weight, a=0, mode="fan_in")

    def forward(self, zx, combine_inner_dims=(1,), combine_index=None, dim_size=None):
        """zx (..., d_latent + d_in), rows view-major (obj, view, point) (resnetfc.py:132-184)."""
        assert zx.size(-1) == self.d_latent + self.d_in
        z, x = zx[..., : self.d_latent], zx[..., self.d_latent:]
        x = self.lin_in(x)
        for i, block in enumerate(self.blocks):
            if i == self.combine_layer and not (len(combine_inner_dims) == 1 and combine_inner_dims[0] == 1):
                x = x.reshape(-1, *combine_inner_dims, x.shape[-1]).mean(dim=1)
            if i < self.combine_layer:
                x = x + self.lin_z[i](z)
            x = block(x)
        return self.lin_out(torch.relu(x))

    @classmethod
    def from_conf(cls, conf, d_in, **kwargs):
        return cls(d_in, n_blocks=conf.get_int("n_blocks", 5), d_hidden=conf.get_int("d_hidden", 128),
                   beta=conf.get_float("beta", 0.0), combine_layer=conf.get_int("combine_layer", 1000),
                   combine_type=conf.get_string("combine_type", "average"),
                   use_spade=conf.get_bool("use_spade", False), **kwargs)

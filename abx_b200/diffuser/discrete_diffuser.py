"""Uniform-rate CTMC over the 20 residue types — API of the reference's diffuser/discrete_diffuser.py.

The reverse tau-leap runs in the C-ABI kernels (`abx_seq_reverse_rates` + the jump application inside
`abx_se3_reverse_step`); the closed-form transition matrix and the once-per-sample forward marginal
are torch ops.
"""
import torch

RESIDUE_NUM = 20        # residue_constants.restype_num


class DiscreteDiffuser:

    def __init__(self, discrete_conf):
        self.discrete_conf = discrete_conf
        self.residue_num = RESIDUE_NUM
        self.rate_const = discrete_conf['rate_const']
        S = self.residue_num
        rate = self.rate_const * (torch.ones(S, S) - torch.eye(S))
        self.rate_matrix = (rate - torch.diag(rate.sum(dim=1))).float()      # discrete_diffuser.py:15-24

    def rate(self, t):
        return self.rate_matrix.to(t.device)[None].expand(t.shape[0], -1, -1)

    def transition(self, t):
        """discrete_diffuser.py:53-67.  The rate matrix r(11^T - S I) has eigenvalues {0, -S r}, so
        V diag(e^{lambda t}) V^T = e^{-S r t} I + (1 - e^{-S r t})/S 11^T (what the kernels evaluate)."""
        S = self.residue_num
        e = torch.exp(-S * self.rate_const * t.float()).view(-1, 1, 1)
        q = (1 - e) / S + e * torch.eye(S, device=t.device)[None]
        return torch.where(q < 1e-8, torch.zeros_like(q), q)

    def sample_ref(self, n_samples, device='cpu'):
        return torch.randint(low=0, high=self.residue_num, size=(n_samples[0], n_samples[1]), device=device)

    def forward_marginal(self, x_0, t):
        """discrete_diffuser.py:72-127 (optimize mode start state; once per sample)."""
        B, D = x_0.shape
        dev = t.device
        qt0 = self.transition(t)
        rate = self.rate(t)
        x_0 = torch.clamp(x_0, min=0, max=self.residue_num - 1)
        bidx = torch.arange(B, device=dev).repeat_interleave(D)
        x_t = torch.distributions.categorical.Categorical(qt0[bidx, x_0.long().flatten(), :]).sample().view(B, D)
        rv = rate[bidx, x_t.long().flatten(), :].clone()
        rv[torch.arange(B * D, device=dev), x_t.long().flatten()] = 0.0
        rv = rv.reshape(B, D, self.residue_num)
        dims = torch.distributions.categorical.Categorical(rv.sum(dim=2)).sample()
        newval = torch.distributions.categorical.Categorical(rv[torch.arange(B, device=dev), dims, :]).sample()
        x_tilde = x_t.clone()
        x_tilde[torch.arange(B, device=dev), dims] = newval
        return x_tilde, qt0, rate

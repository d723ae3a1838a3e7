"""VP-SDE translation diffuser — API of the reference's diffuser/r3_diffuser.py `R3Diffuser`.

The per-step work (score of a predicted x0, reverse step) runs inside the fused C-ABI kernels driven by
`FullDiffuser`; what is left here are the closed-form schedule scalars and the once-per-sample draws,
written as torch ops in the reference's dtype-promotion order (float32 0-d constants against t).
"""
import torch


class R3Diffuser:

    def __init__(self, r3_conf):
        self._r3_conf = r3_conf
        self.min_b = r3_conf['min_b']
        self.max_b = r3_conf['max_b']

    def _scale(self, x):
        return x * torch.tensor(self._r3_conf['coordinate_scaling'], device=x.device)

    def _unscale(self, x):
        return x / torch.tensor(self._r3_conf['coordinate_scaling'], device=x.device)

    def b_t(self, t):
        return torch.tensor(self.min_b, device=t.device) + t * torch.tensor(self.max_b - self.min_b, device=t.device)

    def diffusion_coef(self, t):
        return torch.sqrt(self.b_t(t))[:, None, None]

    def drift_coef(self, x, t):
        return -1 / 2 * self.b_t(t)[:, None, None] * x

    def sample_ref(self, n_samples, device='cpu'):
        return torch.randn(size=(*n_samples, 3), device=device)

    def marginal_b_t(self, t):
        return t * torch.tensor(self.min_b, device=t.device) + (1 / 2) * (t ** 2) * torch.tensor(
            self.max_b - self.min_b, device=t.device)

    def conditional_var(self, t):
        return 1 - torch.exp(-self.marginal_b_t(t))

    def score_scaling(self, t):
        return 1 / torch.sqrt(self.conditional_var(t))

    def calc_trans_0(self, score_t, x_t, t):
        beta_t = self.marginal_b_t(t)[..., None, None]
        return (score_t * (1 - torch.exp(-beta_t)) + x_t) / torch.exp(-1 / 2 * beta_t)

    def score(self, x_t, x_0, t, scale=False):
        """r3_diffuser.py:158-164 (torch ops; the sampler's path is abx_se3_scores)."""
        if scale:
            x_t, x_0 = self._scale(x_t), self._scale(x_0)
        t = t[:, None, None]
        return -(x_t - torch.exp(-1 / 2 * self.marginal_b_t(t)) * x_0) / self.conditional_var(t)

    def distribution(self, x_t, score_t, t, mask, dt):
        x_t = self._scale(x_t)
        g_t = self.diffusion_coef(t)
        f_t = self.drift_coef(x_t, t)
        std = g_t * torch.sqrt(dt)
        mu = x_t - (f_t - g_t ** 2 * score_t) * dt
        if mask is not None:
            mu *= mask[..., None]
        return mu, std

    def forward_marginal(self, x_0, t):
        """r3_diffuser.py:80-105."""
        x_0 = self._scale(x_0)
        log_mean_coeff = (-0.5 * self.marginal_b_t(t)).view(-1, *([1] * (x_0.dim() - 1)))
        mean = torch.exp(log_mean_coeff) * x_0
        std = torch.sqrt(1.0 - torch.exp(2.0 * log_mean_coeff))
        x_t = torch.normal(mean=mean, std=std.expand_as(mean))
        score_t = self.score(x_t, x_0, t)
        return self._unscale(x_t), score_t

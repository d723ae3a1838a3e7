"""IGSO(3) rotation diffuser — API of the reference's diffuser/so3_diffuser.py `SO3Diffuser`.

Tables (pdf / cdf / score norms over the sigma x omega grid) are built by `abx_igso3_build_tables`
(seconds of CPU time in the reference, milliseconds here) or loaded from the reference's own `.npy`
cache, whose directory and file names are kept (so3_diffuser.py:131-148) so either implementation can
reuse the other's cache.  Scores and the geodesic step run in the C-ABI kernels; the rarely used
sampling helpers (prior draw, forward marginal) are a handful of torch ops on the device.
"""
import ctypes
import logging
import os

import numpy as np
import torch

from abx_b200 import lib

SERIES_TERMS = 1000          # L of igso3_expansion / score (so3_diffuser.py:15,72)


def torch_interp(x_new, x, y):
    """abx/utils.py:31-59: batched 1-D linear interpolation (same bin convention)."""
    order = x.argsort(dim=1)
    x = torch.gather(x, -1, order)
    y = torch.gather(y, -1, order)
    b = torch.searchsorted(x.contiguous(), x_new.contiguous(), right=False)     # #{x < x_new}
    b = torch.clamp(b, 0, x.shape[1] - 2)
    x_lo, x_hi = torch.gather(x, -1, b), torch.gather(x, -1, b + 1)
    y_lo, y_hi = torch.gather(y, -1, b), torch.gather(y, -1, b + 1)
    w = (x_new - x_lo) / (x_hi - x_lo + 1e-8)
    w = torch.where(x_new > x[:, -1:], torch.ones_like(w), w)
    w = torch.where(x_new < x[:, :1], torch.zeros_like(w), w)
    return y_lo * (1 - w) + y_hi * w


class SO3Diffuser:

    def __init__(self, so3_conf, consts=None):
        self.schedule = so3_conf['schedule']
        if self.schedule != 'logarithmic':
            raise ValueError(f'Unrecognize schedule {self.schedule}')
        self.min_sigma = so3_conf['min_sigma']
        self.max_sigma = so3_conf['max_sigma']
        self.num_sigma = so3_conf['num_sigma']
        self.num_omega = so3_conf['num_omega']
        self.use_cached_score = so3_conf['use_cached_score']
        self._log = logging.getLogger(__name__)
        self.discrete_omega = torch.linspace(0, np.pi, self.num_omega + 1)[1:]
        self._consts = consts
        self._dev = {}

        rp = lambda x: str(x).replace('.', '_')  # noqa: E731
        cache_dir = os.path.join(
            so3_conf['cache_dir'],
            f'eps_{self.num_sigma}_omega_{self.num_omega}_min_sigma_{rp(self.min_sigma)}_max_sigma_{rp(self.max_sigma)}'
            f'_schedule_{self.schedule}')
        names = [os.path.join(cache_dir, n) for n in ('pdf_vals.npy', 'cdf_vals.npy', 'score_norms.npy')]
        if all(os.path.exists(n) for n in names):
            self._log.info(f'Using cached IGSO3 in {cache_dir}')
            self._pdf, self._cdf, self._score_norms = (torch.from_numpy(np.load(n)) for n in names)
        else:
            self._log.info(f'Computing IGSO3. Saving in {cache_dir}')
            self._pdf, self._cdf, self._score_norms = self.build_tables()
            try:
                os.makedirs(cache_dir, exist_ok=True)
                for n, v in zip(names, (self._pdf, self._cdf, self._score_norms)):
                    np.save(n, v.numpy())
            except OSError as e:                                      # read-only cache dir: keep going
                self._log.warning(f'could not write the IGSO3 cache: {e}')
        # so3_diffuser.py:176-181
        self._score_scaling = torch.sqrt(torch.abs(
            torch.sum(self._score_norms ** 2 * self._pdf, axis=-1) / torch.sum(self._pdf, axis=-1)
        )) / torch.tensor(np.sqrt(3))

    # ---- tables ------------------------------------------------------------------------------------
    def build_tables(self, device=None):
        """so3_diffuser.py:150-166 on the GPU: returns (pdf, cdf, score_norms) as CPU float32 tensors."""
        if not torch.cuda.is_available():
            raise lib.AbxError('building the IGSO(3) tables needs a CUDA device (no CPU fallback); '
                               'point so3.cache_dir at an existing cache to run without one')
        device = torch.device(device or 'cuda')
        sig = self.discrete_sigma.to(device).contiguous()
        om = self.discrete_omega.to(device).contiguous()
        out = [torch.empty(self.num_sigma, self.num_omega, device=device) for _ in range(3)]
        with torch.cuda.device(device):
            lib.check(lib.load().abx_igso3_build_tables(lib.stream(), self.num_sigma, self.num_omega, SERIES_TERMS,
                                                        lib.ptr(sig), lib.ptr(om), *(lib.ptr(o) for o in out)))
        return tuple(o.cpu() for o in out)

    def tables_on(self, device):
        """(score_norms, discrete_sigma, discrete_omega, cdf, score_scaling) resident on `device`."""
        key = str(device)
        if key not in self._dev:
            self._dev[key] = tuple(x.to(device).contiguous() for x in (
                self._score_norms, self.discrete_sigma, self.discrete_omega, self._cdf, self._score_scaling))
        return self._dev[key]

    # ---- schedule (so3_diffuser.py:183-220) ----------------------------------------------------------
    @property
    def discrete_sigma(self):
        return self.sigma(torch.linspace(0.0, 1.0, self.num_sigma))

    def sigma(self, t):
        return torch.log(t * torch.exp(torch.tensor(self.max_sigma)) + (1 - t) * torch.exp(torch.tensor(self.min_sigma)))

    def sigma_idx(self, sigma):
        grid = self.tables_on(sigma.device)[1] if sigma.is_cuda else self.discrete_sigma
        return torch.sum(grid[None, ...] <= sigma[..., None] + 1e-5, -1) - 1

    def t_to_idx_tensor(self, t):
        return self.sigma_idx(self.sigma(t))

    def t_to_idx(self, t):
        return self.t_to_idx_tensor(t).tolist()

    def diffusion_coef(self, t):
        s = self.sigma(t)
        return torch.sqrt(2 * (torch.exp(torch.tensor(self.max_sigma)) - torch.exp(torch.tensor(self.min_sigma))) * s
                          / torch.exp(s))

    # ---- sampling (so3_diffuser.py:222-262) -----------------------------------------------------------
    def sample_igso3(self, t, n_samples):
        device = t.device
        x = torch.rand(n_samples, device=device)
        cdf = self.tables_on(device)[3] if t.is_cuda else self._cdf
        om = (self.tables_on(device)[2] if t.is_cuda else self.discrete_omega)[None].expand(t.shape[0], -1)
        return torch_interp(x, cdf[self.t_to_idx_tensor(t)], om)

    def sample(self, t, n_samples):
        x = torch.randn((*n_samples, 3), device=t.device)
        x /= torch.linalg.norm(x, dim=-1, keepdims=True)
        return x * self.sample_igso3(t, n_samples=n_samples)[..., None]

    def sample_ref(self, n_samples, device='cpu'):
        return self.sample(torch.ones(n_samples[0], device=device), n_samples=n_samples)

    # ---- score (so3_diffuser.py:264-301) ---------------------------------------------------------------
    def score(self, vec, t, eps=1e-6):
        assert eps == 1e-6, 'the kernels implement the reference default eps'
        shape = vec.shape
        B = t.shape[0]
        v = vec.reshape(B, -1, 3).float().contiguous()
        N = v.shape[1]
        tab, sig, om, _, _ = self.tables_on(v.device)
        out = torch.empty_like(v)
        t64 = t.to(torch.float64).contiguous()
        is32 = int(t.dtype != torch.float64)
        L = lib.load()
        with lib.device_guard(v):
            if self.use_cached_score:
                lib.check(L.abx_so3_score_rotvec(lib.stream(), B, N, ctypes.byref(self._consts), lib.ptr(v),
                                                 lib.ptr(t64), is32, lib.ptr(tab), lib.ptr(sig), lib.ptr(om), lib.ptr(out)))
            else:
                lib.check(L.abx_igso3_score_series(lib.stream(), B, N, ctypes.byref(self._consts), lib.ptr(v),
                                                   lib.ptr(t64), is32, lib.ptr(sig), SERIES_TERMS, lib.ptr(out)))
        return out.reshape(shape)

    def score_scaling(self, t):
        tab = self.tables_on(t.device)[4] if t.is_cuda else self._score_scaling
        return tab[self.t_to_idx_tensor(t)]

    def forward_marginal(self, rot_0, t):
        """so3_diffuser.py:303-326."""
        from abx_b200.model import quat_affine as qa
        sampled = self.sample(t, n_samples=rot_0.shape[:-1])
        rot_score = self.score(sampled, t).reshape(rot_0.shape)
        quat_t = qa.quat_multiply(qa.rotvec_to_quat(rot_0), qa.rotvec_to_quat(sampled))
        return qa.quat_to_rotvec(quat_t), rot_score

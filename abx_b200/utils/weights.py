"""Deterministic, platform-independent random weights keyed by parameter NAME.

No checkpoints are available offline, and the reference's own initialisation zeroes every
'final'/'gate' Linear (common_modules.py:17-18), which would let a wrong kernel pass (the IPA output
is multiplied by zeros).  Tests, golden-vector generation and the benchmark therefore all use this
generator: every tensor of a state_dict is drawn from numpy's PCG64 seeded by crc32(name) ^ seed, so
the reference model (in the build container) and the B200 model (on the GPU box) get bit-identical
weights without shipping a 41 MB checkpoint.
"""
import zlib

import numpy as np
import torch


def _draw(name, shape, seed):
    rng = np.random.default_rng([zlib.crc32(name.encode()) & 0xFFFFFFFF, seed & 0xFFFFFFFF])
    return rng.standard_normal(size=tuple(shape), dtype=np.float64)


def seeded_tensor(name, shape, seed=0):
    """Value for parameter `name`: LayerNorm/1-D 'weight' ~ 1 + 0.1 n, biases ~ 0.1 n, matrices and
    embedding tables ~ n / sqrt(shape[1:]), trainable_point_weights ~ softplus^-1(1) + 0.3 n."""
    shape = tuple(shape)
    x = _draw(name, shape, seed)
    leaf = name.rsplit('.', 1)[-1]
    if leaf == 'trainable_point_weights':
        x = np.log(np.e - 1.0) + 0.3 * x
    elif len(shape) == 1:
        x = (1.0 + 0.1 * x) if leaf == 'weight' else 0.1 * x
    else:
        fan_in = int(np.prod(shape[1:]))
        x = x / np.sqrt(fan_in)
        if 'affine_update' in name:
            x = 0.1 * x        # keep per-layer frame updates ~1 A / ~0.1 rad, as in a trained model
    return torch.from_numpy(x.astype(np.float32))


def seeded_state_dict(shapes, seed=0):
    """shapes: mapping name -> shape (e.g. {k: v.shape for k, v in model.state_dict().items()})."""
    return {k: seeded_tensor(k, s, seed) for k, s in shapes.items()}


def load_seeded_(module, seed=0):
    sd = module.state_dict()
    new = seeded_state_dict({k: v.shape for k, v in sd.items()}, seed)
    new = {k: v.to(dtype=sd[k].dtype) for k, v in new.items()}
    module.load_state_dict(new, strict=True)
    return module


def np_randn(seed, *shape):
    """Platform-independent standard-normal float32 tensor (numpy PCG64), for test inputs that are too
    large to store as fixtures."""
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(size=shape).astype(np.float32))

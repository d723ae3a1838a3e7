"""abx_b200 — B200-native reverse-diffusion sampling hot path of AbX."""
__version__ = '0.1.0'

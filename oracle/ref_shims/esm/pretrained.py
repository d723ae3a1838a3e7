"""Stub: ESM2 weights are unavailable offline; the oracle runs with esm.enabled=false."""


def load_model_and_alphabet_local(path):
    raise RuntimeError('ESM2 is not available in this environment (esm.enabled must be false)')

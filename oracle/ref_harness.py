"""Import harness for the read-only reference checkout (/root/reference).

TEST INFRASTRUCTURE ONLY.  This module exists so that `oracle/make_golden.py` can import the
reference's own hot-path modules *in the build container* and dump golden vectors.  Nothing in
tests/, bench.py or the product package imports it at run time: /root/reference does not exist on
the GPU box.

The reference needs four tiny stand-ins: `tree`, `ml_collections`, `esm.pretrained`
(in oracle/ref_shims/) and import-time stubs for Bio / anarci / pyrosetta (below).
"""
import importlib.abc
import importlib.machinery
import json
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('ABX_REFERENCE_ROOT', '/root/reference')
_HERE = os.path.dirname(os.path.abspath(__file__))


class _StubModule(types.ModuleType):
    """Module whose every attribute is another stub (import-time names only)."""
    __path__ = []

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        sub = _StubModule(f'{self.__name__}.{name}')
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        raise RuntimeError(f'{self.__name__} is a stub (dependency not installed)')


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ('Bio', 'anarci', 'pyrosetta')

    def find_spec(self, fullname, path, target=None):
        if fullname.split('.')[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Make `import abx...` / `import diffuser...` resolve to the reference checkout."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f'reference checkout not found at {REFERENCE_ROOT}')
    sys.path.insert(0, os.path.join(_HERE, 'ref_shims'))
    sys.path.insert(0, REFERENCE_ROOT)
    sys.meta_path.append(_StubFinder())
    _installed = True


def load_config(esm_enabled=False, use_cached_score=True, cache_dir=None):
    """config/config_model.json as ConfigDict with the inference-time overrides
    (inference.py:93-99) and ESM disabled (weights unavailable offline)."""
    install()
    import ml_collections
    with open(os.path.join(REFERENCE_ROOT, 'config', 'config_model.json')) as f:
        raw = json.load(f)
    raw['model']['embeddings_and_seqformer']['esm']['enabled'] = esm_enabled
    raw['diffuser']['so3']['use_cached_score'] = use_cached_score
    if cache_dir is not None:
        raw['diffuser']['so3']['cache_dir'] = cache_dir
    return ml_collections.ConfigDict(raw), raw

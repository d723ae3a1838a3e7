"""CPU checks of the drop-in boundary: the C-ABI library loads and exports exactly what include/abx_b200.h
declares, and the ctypes signatures mirror the header."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'abx_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    out = {}
    for m in re.finditer(r'\b([A-Za-z_][\w\s\*]*?)\b(abx_\w+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ('', 'void') else len([a for a in args.split(',') if a.strip()])
        out[m.group(2)] = n
    return out


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from abx_b200 import lib
    L = lib.load()
    decl = header_functions()
    assert len(decl) >= 15
    for name in decl:
        assert hasattr(L, name), f'{name} declared in include/abx_b200.h but not exported'
    assert L.abx_version() == 1


def test_ctypes_signatures_mirror_the_header():
    from abx_b200 import lib
    decl = header_functions()
    assert set(decl) == set(lib.SIGNATURES), set(decl) ^ set(lib.SIGNATURES)
    for name, n in decl.items():
        assert len(lib.SIGNATURES[name][1]) == n, name


def test_struct_layouts():
    from abx_b200 import lib
    assert ctypes.sizeof(lib.DiffuserConsts) == 7 * 8 + 2 * 4
    assert ctypes.sizeof(lib.IpaWeights) == 15 * 8


def test_errors_are_reported_not_thrown():
    """Argument validation happens before any CUDA call, so it can be exercised without a GPU."""
    from abx_b200 import lib
    L = lib.load()
    rc = L.abx_linear_f32(None, 4, 4, 3, None, 3, None, None, None, 0, None, 4)
    assert rc == 1 and b'abx_linear_f32' in L.abx_last_error()
    rc = L.abx_ipa_forward(None, 0, 5, None, None, None, None, None, None, None, None, None, None, 0)
    assert rc == 1
    assert L.abx_ipa_workspace_bytes(2, 350) > 2 * 12 * 350 * 350 * 4

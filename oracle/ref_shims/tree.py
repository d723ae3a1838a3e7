"""Stand-in for dm-tree, only so the read-only reference imports in this container.
Used solely by oracle/make_golden.py (fixture generation); never at test/bench run time."""


def map_structure(fn, *structures):
    first = structures[0]
    if isinstance(first, (list, tuple)):
        out = [map_structure(fn, *xs) for xs in zip(*structures)]
        return type(first)(out) if not hasattr(first, '_fields') else type(first)(*out)
    if isinstance(first, dict):
        return {k: map_structure(fn, *[s[k] for s in structures]) for k in first}
    return fn(*structures)

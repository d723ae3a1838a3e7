"""Materialise oracle/_ref/ from the read-only reference checkout (the Python analogue of compiling a C reference from its
own sources): the hot-path packages of /root/reference (abx/, diffuser/, config/, inference.py) are copied UNMODIFIED into
oracle/_ref/ — git-ignored, so no reference source enters the history, but it travels to the GPU box with the snapshot —
and the reference's own FullDiffuser builds its IGSO(3) cache into oracle/_ref/igso3_cache/ once (72 s of CPU work that
would otherwise be paid on the GPU box).  `bench.py --impl reference` and `cpu_baseline` then time the reference's own
modules (oracle/ref_runner.py, cpu_baseline.kind = "reference"); without oracle/_ref they fall back to the oracle port.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under abx_b200/ imports oracle/.

    python oracle/build_ref.py [--force]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.environ.get('ABX_REFERENCE_SRC', '/root/reference')
DST = os.path.join(HERE, '_ref')
PARTS = ('abx', 'diffuser', 'config', 'inference.py')


def build_ref(force=False):
    """-> path of oracle/_ref, or None when the reference checkout is not present (GPU box: uses what travelled)."""
    stamp = os.path.join(DST, '.complete')
    if os.path.exists(stamp) and not force:
        return DST
    if not os.path.isdir(SRC):
        return DST if os.path.exists(stamp) else None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for part in PARTS:
        s, d = os.path.join(SRC, part), os.path.join(DST, part)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        else:
            shutil.copy2(s, d)
    # the reference's own IGSO(3) tables (so3_diffuser.py:150-181), built by the reference's code
    sys.path.insert(0, ROOT)
    os.environ['ABX_REFERENCE_ROOT'] = DST
    from oracle import ref_harness
    ref_harness.REFERENCE_ROOT = DST
    ref_harness.install()
    cache = os.path.join(DST, 'igso3_cache') + os.sep
    os.makedirs(cache, exist_ok=True)
    from diffuser.full_diffuser import FullDiffuser
    cfg, _ = ref_harness.load_config(cache_dir=cache)
    FullDiffuser.get(cfg.diffuser)
    open(stamp, 'w').write('ok\n')
    return DST


if __name__ == '__main__':
    print(build_ref(force='--force' in sys.argv))

"""In-tree build of the sm_100a C-ABI library (abx_b200/csrc/libabx_b200.so) with nvcc.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with the
gpurun snapshot.  `python -m abx_b200.build` rebuilds unconditionally.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
LIB = os.path.join(CSRC, 'libabx_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
         '-Xptxas', '-v', '-I', INCLUDE, '-I', CSRC] + \
        ([f'-DABX_GEMM_PROFILE={os.environ["ABX_GEMM_PROFILE"]}'] if os.environ.get('ABX_GEMM_PROFILE') else []) + \
        ([f'-DABX_ATTN_PROFILE={os.environ["ABX_ATTN_PROFILE"]}'] if os.environ.get('ABX_ATTN_PROFILE') else [])


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(INCLUDE, '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build_extension(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library; returns its path."""
    if not force and not _stale():
        return LIB
    objs, procs = [], []
    os.makedirs(os.path.join(CSRC, 'build'), exist_ok=True)
    for src in sources():
        obj = os.path.join(CSRC, 'build', os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        procs.append((src, subprocess.Popen([NVCC, *FLAGS, '-c', src, '-o', obj], stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f'nvcc failed on {src}')
    with open(os.path.join(CSRC, 'build', 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    subprocess.check_call([NVCC, '-shared', '-o', LIB, *objs, '-cudart', 'static'])
    return LIB


if __name__ == '__main__':
    print(build_extension(force=True, verbose='-v' in sys.argv))

"""Compile csrc/*.cu into dmf/libdmf_b200.so for sm_100a with nvcc (in-tree, so the .so travels to
the GPU box with the repo snapshot)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), 'csrc')
LIB = os.path.join(HERE, 'libdmf_b200.so')
OBJ = os.path.join(HERE, '_obj')
SOURCES = ['scene.cu', 'metrics.cu', 'ihs.cu', 'net.cu', 'dense.cu', 'train.cu']
FLAGS = ['-std=c++17', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xcompiler', '-fPIC']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), 'include', 'dmf_b200.h'))
    nvcc = _nvcc()
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + '.o')
        if force or _stale(obj, [src] + headers):
            jobs.append([nvcc] + FLAGS + ['-c', src, '-o', obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed: %s\n%s\n%s' % (' '.join(cmd), r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, s[:-3] + '.o') for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        run([nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))

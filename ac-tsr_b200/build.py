"""Build libacsr.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python ac-tsr_b200/build.py [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(CSRC, 'libacsr.so')
STAMP = os.path.join(CSRC, '.libacsr.stamp')
SOURCES = ['api.cu', 'rowwise.cu', 'attn_fwd.cu', 'attn_bwd.cu', 'logits_tc.cu', 'topk_merge.cu', 'linear.cu', 'linear_tok.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError('nvcc not found')


def _digest():
    h = hashlib.sha256()
    root = os.path.dirname(HERE)
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cu', '.cuh', '.h'))]
    files.append(os.path.join(root, 'include', 'acsr.h'))
    for f in files:
        with open(f, 'rb') as fh:
            h.update(f.encode() + b'\0' + fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile if sources changed; returns the path of the shared library."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(CSRC, s[:-3] + '.o')
        cmd = [_nvcc()] + NVCC_FLAGS + ['-c', os.path.join(CSRC, s), '-o', o]
        procs.append((s, o, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, o, p in procs:
        out, _ = p.communicate()
        log.append('== %s ==\n%s' % (s, out))
        if p.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (s, out))
        objs.append(o)
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout)
    with open(os.path.join(CSRC, 'build.log'), 'w') as fh:
        fh.write('\n'.join(log))
    with open(STAMP, 'w') as fh:
        fh.write(dig)
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))

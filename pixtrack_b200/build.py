"""Builds pixtrack_b200/libpixtrack_b200.so (sm_100a) in-tree with nvcc.

    python -m pixtrack_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
nvcc cross-compiles without a GPU, so this also runs in the CPU-only
authoring container (the driver's "does it build" check calls it through
__graft_entry__.build()).
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(HERE, '_build')
LIB = os.path.join(HERE, 'libpixtrack_b200.so')
SOURCES = ('ptk_api.cu', 'ptk_lm.cu', 'ptk_sample.cu', 'ptk_sample_ref.cu', 'ptk_conv.cu', 'ptk_unet.cu', 'ptk_nerf.cu', 'ptk_mask.cu', 'ptk_overlay.cu')
# per-file extra flags: the NeRF marcher is compiled without FMA contraction so that its geometric decisions
# (voxel stepping, occupancy tests) are the same IEEE operation sequence as the numpy oracle
EXTRA_FLAGS = {'ptk_nerf.cu': ['-fmad=false']}
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v', '-I', os.path.join(ROOT, 'include'), '-I', CSRC]


def _nvcc() -> str:
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; cannot build libpixtrack_b200.so')
    return exe


def _digest() -> str:
    h = hashlib.sha256((' '.join(NVCC_FLAGS) + repr(sorted(EXTRA_FLAGS.items()))).encode())
    names = sorted(os.listdir(CSRC)) + ['../../include/pixtrack_b200.h']
    for n in names:
        with open(os.path.join(CSRC, n), 'rb') as f:
            h.update(n.encode() + f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, 'digest.txt')
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()
    objs = []
    log = []
    for src in SOURCES:
        obj = os.path.join(BUILD, src.replace('.cu', '.o'))
        cmd = [nvcc, *NVCC_FLAGS, *EXTRA_FLAGS.get(src, []), '-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f'nvcc failed on {src}')
        objs.append(obj)
    r = subprocess.run([nvcc, '-shared', '-o', LIB, *objs, '-gencode', 'arch=compute_100a,code=sm_100a'],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('link failed')
    with open(os.path.join(BUILD, 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    with open(stamp, 'w') as f:
        f.write(dig)
    if verbose:
        sys.stderr.write('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))

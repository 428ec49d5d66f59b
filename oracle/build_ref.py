"""Test infrastructure: builds oracle/_ref/ngp_host from oracle/ngp_ref/ngp_host.cu against the instant-ngp headers
IN PLACE under /root/reference (authoring container only; nothing is copied, outputs go to oracle/_ref/ which is
git-ignored).  The reference's own build (cmake, GUI / Vulkan / pybind dependencies) is not used: the harness needs
only header-only code (Eigen, tinylogger, fmt, tiny-cuda-nn's common.h).  nvcc is used as the host compiler driver so
that the headers' __host__ __device__ annotations parse; -arch=sm_80 only satisfies tiny-cuda-nn's static assert --
no device code is run.

    python oracle/build_ref.py          # -> oracle/_ref/ngp_host
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/instant-ngp'


def build(verbose: bool = True) -> str:
    """Returns the path of the binary, or '' when the reference tree is absent (e.g. on the GPU box)."""
    if not os.path.isdir(os.path.join(REF, 'include', 'neural-graphics-primitives')):
        return ''
    out_dir = os.path.join(HERE, '_ref')
    os.makedirs(out_dir, exist_ok=True)
    src, out = os.path.join(HERE, 'ngp_ref', 'ngp_host.cu'), os.path.join(out_dir, 'ngp_host')
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    tcnn = os.path.join(REF, 'dependencies', 'tiny-cuda-nn')
    inc = [os.path.join(REF, 'include'), os.path.join(REF, 'dependencies'), os.path.join(REF, 'dependencies', 'tinylogger'),
           os.path.join(REF, 'dependencies', 'eigen'), os.path.join(REF, 'dependencies', 'filesystem'),
           os.path.join(tcnn, 'include'), os.path.join(tcnn, 'dependencies'), os.path.join(tcnn, 'dependencies', 'fmt', 'include')]
    cmd = ['nvcc', '-std=c++14', '-arch=sm_80', '--extended-lambda', '--expt-relaxed-constexpr', '-w', '-x', 'cu',
           '-DFMT_HEADER_ONLY', '-DTCNN_MIN_GPU_ARCH=80', '-DNGP_VERSION="ref"'] + [f'-I{i}' for i in inc] + [src, '-o', out]
    if verbose:
        print(' '.join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return out


if __name__ == '__main__':
    print(build())

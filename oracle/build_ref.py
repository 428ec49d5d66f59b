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


# testbed_nerf.cu keeps its marching helpers local to the translation unit (and marks most of them __device__);
# the whole file needs the GUI / Testbed headers.  These definitions are therefore lifted out AT BUILD TIME into a
# temporary include (deleted after the compile; never committed), with `__device__` widened to
# `__host__ __device__` so that the harness can call them on the CPU.  The function bodies are untouched.
HELPERS = ['NERF_RENDERING_NEAR_DISTANCE', 'NERF_STEPS', 'NERF_CASCADES', 'SQRT3', 'STEPSIZE', 'MIN_CONE_STEPSIZE',
           'MAX_CONE_STEPSIZE', 'grid_mip_offset', 'calc_dt', 'distance_to_next_voxel', 'advance_to_next_voxel',
           'warp_position', 'unwarp_position', 'warp_direction', 'warp_dt', 'unwarp_dt', 'cascaded_grid_idx_at',
           'density_grid_occupied_at', 'mip_from_pos', 'mip_from_dt',
           # compositing: the activations (all overloads) and the kernel itself, lifted as a host function
           'network_to_rgb*', 'network_to_density', 'network_to_density_derivative', 'composite_kernel_nerf',
           # start of a ray and end of a sample pass: ray init, jittered first advance, shade
           'calc_cone_angle', 'advance_pos_nerf', 'init_rays_with_payload_kernel_nerf', 'shade_kernel_nerf',
           'compact_kernel_nerf', 'generate_next_nerf_network_inputs']
RENDER_BUFFER_HELPERS = ['accumulate_kernel', 'tonemap*', 'tonemap_kernel']   # src/render_buffer.cu:236-345,542-569


TCNN_HELPERS = ['fast_hash', 'grid_index', 'kernel_grid']   # tiny-cuda-nn/include/tiny-cuda-nn/encodings/grid.h:82-116,135-340
TCNN_DEVICE_FUNS = ['identity_fun', 'identity_derivative', 'smoothstep', 'smoothstep_derivative']   # common_device.h:379-397
TCNN_SH = ['kernel_sh']                                      # encodings/spherical_harmonics.h:46-150
# The two kernels are lifted as ordinary host functions (`__global__` -> `__host__`); the harness supplies
# threadIdx / blockIdx / blockDim as host variables through macros, so their bodies stay untouched as well.


def lift_helpers(cu_path: str, names=None) -> str:
    import re
    lines = open(cu_path).read().split('\n')
    out = []
    for name in (names or HELPERS):
        every = name.endswith('*')
        name = name.rstrip('*')
        pat = re.compile(r'^(inline |static )?(constexpr )?(__host__ )?(__device__ |__global__ )[\w:<>, &\*]*\b' + name + r'\(')
        cands = [i for i, ln in enumerate(lines) if pat.match(ln)]
        if name == 'pos_fract':            # three overloads: the (pos, pos_derivative, pos_grid) one is what kernel_grid calls
            cands = [i for i in cands if 'pos_derivative' in lines[i] and 'pos_2nd_derivative' not in lines[i]]
        for start in (cands if every else cands[:1]):
            out.extend(_one_definition(lines, start))
    return '\n'.join(out)


def lift_member(header_path: str, name: str) -> str:
    """The definition of member function `name` (first one with a body) out of a class in a header, verbatim."""
    import re
    lines = open(header_path).read().split('\n')
    pat = re.compile(r'^\s*(virtual\s+)?[\w:<>\*& ]+\b' + name + r'\(.*\{\s*$')
    start = next(i for i, ln in enumerate(lines) if pat.match(ln))
    depth, i = 0, start
    while True:
        depth += lines[i].count('{') - lines[i].count('}')
        if depth == 0:
            break
        i += 1
    return '\n'.join(lines[start:i + 1]) + '\n'


def _one_definition(lines, start):
    out = []
    if lines[start - 1].startswith('template'):
        out.append(lines[start - 1])
    depth, i = 0, start
    while True:
        depth += lines[i].count('{') - lines[i].count('}')
        if depth == 0 and '{' in ''.join(lines[start:i + 1]):
            break
        i += 1
    body = lines[start:i + 1]
    if '__global__' in body[0]:
        body[0] = body[0].replace('__global__', '__host__', 1)
    elif '__host__' not in body[0]:
        body[0] = body[0].replace('__device__', '__host__ __device__', 1)
    out.extend(body + [''])
    return out


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
    tmp_inc = os.path.join(out_dir, 'tmp_include')
    os.makedirs(tmp_inc, exist_ok=True)
    lifted = os.path.join(tmp_inc, 'testbed_nerf_helpers.inc')
    with open(lifted, 'w') as f:
        f.write(lift_helpers(os.path.join(REF, 'src', 'testbed_nerf.cu')))
        f.write('\n')
        f.write(lift_helpers(os.path.join(REF, 'src', 'render_buffer.cu'), RENDER_BUFFER_HELPERS))
    lifted2 = os.path.join(tmp_inc, 'tcnn_grid_helpers.inc')
    with open(lifted2, 'w') as f:
        tinc = os.path.join(tcnn, 'include', 'tiny-cuda-nn')
        f.write(lift_helpers(os.path.join(tinc, 'common_device.h'), TCNN_DEVICE_FUNS + ['pos_fract']))
        f.write('\n')
        f.write(lift_helpers(os.path.join(tinc, 'encodings', 'grid.h'), TCNN_HELPERS))
        f.write('\n')
        f.write(lift_helpers(os.path.join(tinc, 'encodings', 'spherical_harmonics.h'), TCNN_SH))
    lifted3 = os.path.join(tmp_inc, 'nerf_network_set_params.inc')
    with open(lifted3, 'w') as f:
        f.write(lift_member(os.path.join(REF, 'include', 'neural-graphics-primitives', 'nerf_network.h'), 'set_params'))
    cmd.insert(-3, f'-I{tmp_inc}')
    if verbose:
        print(' '.join(cmd), file=sys.stderr)
    try:
        subprocess.run(cmd, check=True)
    finally:
        os.remove(lifted)
        os.remove(lifted2)
        os.remove(lifted3)
        os.rmdir(tmp_inc)
    return out


if __name__ == '__main__':
    print(build())

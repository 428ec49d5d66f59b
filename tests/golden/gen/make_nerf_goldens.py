"""Generate tests/golden/nerf_host.json from the UNMODIFIED instant-ngp headers.

Run in the authoring container only (needs /root/reference and nvcc):

    python tests/golden/gen/make_nerf_goldens.py

oracle/build_ref.py compiles oracle/ngp_ref/ngp_host.cu against the reference headers in place; the binary calls the
reference's host-callable functions (jitter sequence, colour transfer, focal length, camera-matrix conversion, ray
generation, box intersection) and its JSON output is stored verbatim.  tests/test_nerf_oracle.py checks
oracle/nerf.py against it.
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..', '..'))
sys.path.insert(0, ROOT)
from oracle.build_ref import build  # noqa: E402

if __name__ == '__main__':
    exe = build()
    assert exe, 'reference tree not found'
    data = json.loads(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    path = os.path.join(HERE, '..', 'nerf_host.json')
    with open(path, 'w') as f:
        json.dump(data, f)
    hits = sum(1 for s in data['rays']['samples'] if s['t_box1'][0] < 1e30)
    print('wrote', os.path.abspath(path), os.path.getsize(path), 'bytes;', len(data['rays']['samples']), 'rays,', hits, 'hit box 1')

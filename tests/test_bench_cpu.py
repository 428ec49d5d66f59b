"""CPU: the reference arm of bench.py (`--impl reference`: the oracle port on the host cores) runs without a GPU and
prints one JSON line with the keys the driver reads; the workload table is consistent."""
import json
import os
import subprocess

import numpy as np
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                        '--workload', 'c3'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['metric'] == 'tracked frames/sec (LM-to-convergence)' and d['value'] > 0 and d['steps'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] == (os.cpu_count() or 1)
    assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['config']['workload'].startswith('C3 stand-in') and d['vs_baseline'] is None and d['data'] == 'synthetic'


def test_workload_table():
    sys.path.insert(0, ROOT)
    import bench
    assert sorted(bench.WORKLOADS) == ['c2', 'c3', 'c4', 'c5']
    assert bench.WORKLOADS['c2']['n_points'] == 5000 and bench.WORKLOADS['c2']['n_views'] == 8
    assert bench.WORKLOADS['c4']['n_points'] == 20000 and bench.WORKLOADS['c4']['n_views'] == 16
    assert bench.WORKLOADS['c4']['stop']['num_iters'] == 30 and bench.WORKLOADS['c4']['stop']['dt_stop'] == 0.0
    assert bench.WORKLOADS['c5']['nerf'] and not bench.WORKLOADS['c2']['nerf']
    seq = bench.make_sequence(dict(bench.WORKLOADS['c3'], n_points=50), 1)
    cq, cr = seq['cam_q'], seq['cam_r']
    assert np.allclose([float(x) for x in cq[:6]], [640.0, 480.0, 1066.778, 1067.487, 319.5, 239.5], rtol=1e-6)   # YCB intrinsics, c forced
    assert [float(x) for x in cr[:2]] == [192.0, 144.0] and abs(float(cr[2]) - 0.3 * 1066.778) < 1e-3
    assert tuple(seq['frames'][0]['img_q'].shape) == (480, 640, 3) and tuple(seq['frames'][0]['img_r'].shape) == (144, 192, 3)

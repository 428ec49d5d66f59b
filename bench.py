#!/usr/bin/env python
"""bench.py -- tracked frames/s of the PixTrack pose-refinement hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1] restated on synthetic data, SURVEY 8d "C2"):
one step = one tracked frame = coarse-to-fine LM-to-convergence (reference
stop criteria, num_iters=150) over the 3-level feature pyramid of a 1024x576
query (32x576x1024 / 128x144x256 / 128x36x64), N=5000 3D points, B=8
reference views solved together.  N>1: one process per GPU, independent
frames sharded over ranks (weak scaling), one NCCL all_gather of the poses at
the end.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pixtrack_b200 import synthetic as syn  # noqa: E402

N_POINTS, N_VIEWS, RING = 5000, 8, 3
WORKLOAD = ('C2: 1024x576 query pyramid (32x576x1024,128x144x256,128x36x64 fp32), N=5000 points, B=8 reference '
            'views, 3 levels coarse-to-fine, LM to convergence (num_iters=150, stop 1e-4/5e-3/5e-2), damping const=0')
STOP = dict(num_iters=150, grad_stop=1e-4, dt_stop=5e-3, dR_stop=5e-2)


def lam0():
    return 10.0 ** (-6.0 + torch.sigmoid(torch.zeros(6)) * 11.0)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace('.', '').isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        mx = float(self.rows[0][2]) if self.rows and len(self.rows[0]) >= 8 else None
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def host_frames(rank):
    return [syn.frame_problem(seed=100 + 17 * rank + i, N=N_POINTS, B=N_VIEWS) for i in range(RING)]


# ------------------------------------------------------------------------------------------------
# reference arm: the oracle port of the reference's CPU path (the reference is Python and is not
# present on the GPU box; oracle/lm.py restates it with the same torch CPU primitives)
# ------------------------------------------------------------------------------------------------
def cpu_refine_view(fr, view, threads):
    from oracle import lm
    torch.set_num_threads(threads)
    R, t = fr['T_init'][view, :9].reshape(3, 3), fr['T_init'][view, 9:]
    iters = []
    for lv in (2, 1, 0):
        out = lm.lm_run(fr['p3d'], fr['F_ref'][lv][view], fr['F_q'][lv], R, t, fr['cam'][lv],
                        fr['W_ref'][lv][view][:, None], fr['W_q'][lv], lam=lam0(), **STOP)
        iters.append(out['n_iters'])
        if out['failed']:
            break
        R, t = out['R'], out['t']
    return iters


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_grad_enabled(False)
    threads = os.cpu_count() or 1
    frames = [syn.frame_problem(seed=100 + i, N=N_POINTS, B=N_VIEWS) for i in range(1)]
    for _ in range(min(args.warmup, 1)):
        cpu_refine_view(frames[0], 0, threads)
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_refine_view(frames[0], i % N_VIEWS, threads)
    dt = time.perf_counter() - t0
    fps = args.steps / dt / N_VIEWS          # a frame is 8 views
    line = {
        'impl': 'reference', 'metric': 'tracked frames/sec (LM-to-convergence)', 'value': fps, 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps * N_VIEWS,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD},
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{args.steps} steps, each = 1 of the 8 views of one frame (3 levels, to convergence); '
                                   'frames/s = views/s / 8'},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from pixtrack_b200 import _lib
    from pixtrack_b200.optimizer import query_map_to_hwc
    from pixtrack_b200.refiner import refine_levels_batched

    torch.set_grad_enabled(False)
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lam = lam0().to(dev)
    frames = host_frames(rank)

    # host (pinned) and device copies of the per-frame inputs
    host, devf = [], []
    for fr in frames:
        h = dict(fq=[f.permute(1, 2, 0).contiguous().pin_memory() for f in fr['F_q']],
                 wq=[w[0].contiguous().pin_memory() for w in fr['W_q']])
        d = dict(fq=[x.to(dev) for x in h['fq']], wq=[x.to(dev) for x in h['wq']],
                 cam=[c.to(dev) for c in fr['cam']], F_ref=[x.to(dev) for x in fr['F_ref']],
                 W_ref=[x.to(dev) for x in fr['W_ref']], p3d=fr['p3d'].to(dev), T0=fr['T_init'].to(dev))
        host.append(h)
        devf.append(d)
    ring_bytes = sum(x.numel() * 4 for x in host[0]['fq'] + host[0]['wq']) * RING

    launches = [0]

    def step(d, fq=None, wq=None):
        launches[0] += 3
        return refine_levels_batched(fq or d['fq'], wq or d['wq'], d['cam'], d['F_ref'], d['W_ref'], d['p3d'], d['T0'],
                                     [lam] * 3, **STOP)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing --------------------------------------------------------------
    for i in range(args.warmup):
        step(devf[i % RING])
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    outs = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches[0] = 0
    ev0.record()
    for i in range(args.steps):
        outs.append(step(devf[i % RING]))
    ev1.record()
    barrier()
    n_launch = launches[0]
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    clk = clocks.stop() if rank == 0 else None
    _lib.device_status(local_rank)

    # ---- per-launch timing of the LM kernel for the roofline (same inputs, events around each launch)
    from pixtrack_b200.optimizer import lm_run_batched
    per_level = {}
    for i in range(min(args.steps, 2 * RING)):
        d = devf[i % RING]
        T, skip = d['T0'], None
        for lv in (2, 1, 0):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            T, failed, n_it, log = lm_run_batched(d['p3d'], d['F_ref'][lv], d['fq'][lv], T, d['cam'][lv], lam,
                                                  d['W_ref'][lv], d['wq'][lv], None, skip, **STOP)
            b.record()
            torch.cuda.synchronize()
            skip = failed
            C = d['fq'][lv].shape[-1]
            nv = 0.0
            lg = log.cpu()
            for v in range(N_VIEWS):
                nv += float(lg[v, :int(n_it[v]), 1].sum())
            rec = per_level.setdefault(lv, dict(ms=0.0, bytes=0.0, n=0, iters=0))
            rec['ms'] += a.elapsed_time(b)
            rec['bytes'] += nv * (52 * C + 32)
            rec['n'] += 1
            rec['iters'] += int(n_it.max())
    dom = max(per_level, key=lambda k: per_level[k]['ms'])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    ach = per_level[dom]['bytes'] / (per_level[dom]['ms'] * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': f'lm_kernel level {dom}', 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                'frac': ach / peak, 'traffic': None,
                'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650 GB/s',
                'per_level': {str(k): {'ms_per_launch': v['ms'] / v['n'], 'GBps': v['bytes'] / (v['ms'] * 1e-3) / 1e9,
                                       'max_iters_per_launch': v['iters'] / v['n']} for k, v in per_level.items()}}

    # ---- end to end: host buffers in, poses out ------------------------------------------------
    def e2e_step(i):
        h, d = host[i % RING], devf[i % RING]
        fq = [x.to(dev, non_blocking=True) for x in h['fq']]
        wq = [x.to(dev, non_blocking=True) for x in h['wq']]
        out = step(d, fq, wq)
        return out['T'].cpu(), out['failed'].cpu()
    for i in range(max(1, args.warmup // 2)):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    h2d = sum(x.numel() * 4 for x in host[0]['fq'] + host[0]['wq'])
    d2h = N_VIEWS * 13

    # ---- final gather of per-frame results (the only collective on this path) -------------------
    res = torch.stack([torch.cat([o['T'], o['failed'].float()[:, None]], 1) for o in outs[-RING:]])
    if world > 1:
        gathered = [torch.empty_like(res) for _ in range(world)]
        dist.all_gather(gathered, res)
    iters = [[int(x.max()) for x in o['n_iters']] for o in outs[-RING:]]
    ok = all(not bool(o['failed'].any()) for o in outs)

    if rank == 0:
        cpu = None
        if world == 1:
            th = os.cpu_count() or 1
            t0 = time.perf_counter()
            n_v = 0
            while n_v < 2 and time.perf_counter() - t0 < 25:
                cpu_refine_view(frames[0], n_v, th)
                n_v += 1
            dt = time.perf_counter() - t0
            cpu = {'value': n_v / dt / N_VIEWS, 'unit': 'frames/s', 'cores': th, 'kind': 'port',
                   'sample': f'{n_v} of the 8 views of one C2 frame (3 levels to convergence) via oracle/lm.py, '
                             f'{dt:.1f} s; frames/s = views/s / 8'}
        fps = args.steps * world / (ms_total * 1e-3)
        line = {
            'metric': 'tracked frames/sec (LM-to-convergence)', 'value': fps, 'unit': 'frames/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_total / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'l2': f'inputs larger than L2: ring of {RING} frames, {ring_bytes / 1e6:.0f} MB of maps',
                       'stage': 'LM only (pyramids synthetic; extractor and NeRF render not in the timed step yet)',
                       'lm_iters_last_frames_coarse_to_fine': iters, 'all_converged_without_failure': ok},
            'e2e': {'value': args.steps * world / float(e2e_s), 'unit': 'frames/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h},
            'gpu_launches': n_launch, 'clocks': clk, 'roofline': roofline, 'cpu_baseline': cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()

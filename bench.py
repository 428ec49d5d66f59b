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
def lm_stress(dev, lam, peak):
    """LM kernel alone at BASELINE config-4 scale (level 1: C=128, 144x256, N=20000, B=16, 30 fixed
    iterations): the regime where a roofline fraction is meaningful."""
    from pixtrack_b200.optimizer import LmLaunch, query_map_to_hwc
    p = syn.level_problem(seed=9, N=20000, C=128, H=144, W=256, B=16, noise=0.02, rot_deg=0.5, trans=0.005)
    T0 = torch.cat([p['R0'].reshape(16, 9), p['t0']], 1).to(dev)
    L = LmLaunch(p['p3d'].to(dev), p['F_ref'].to(dev), query_map_to_hwc(p['F_q'].to(dev)), T0, p['cam'].to(dev), lam,
                 p['W_ref'].reshape(16, -1).to(dev), p['W_q'].to(dev), num_iters=30, grad_stop=0.0, dt_stop=0.0,
                 dR_stop=0.0)
    for _ in range(3):
        L.launch()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    a.record()
    for _ in range(reps):
        L.launch()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    nv = float(L.log[:, :, 1].sum())
    byts = nv * (52 * 128 + 32)
    g, ng = L.plan()
    return {'workload': 'C4 level 1: C=128 144x256, N=20000, B=16, 30 fixed iterations', 'ms_per_launch': ms,
            'us_per_iteration': 1e3 * ms / 30, 'achieved': byts / (ms * 1e-3) / 1e9, 'unit': 'GB/s',
            'frac': byts / (ms * 1e-3) / 1e9 / peak, 'ctas_per_problem': g, 'problems_in_flight': ng,
            'note': 'algorithmic bytes (52C+32 per valid point per iteration); the 19 MB map stays L2-resident'}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from pixtrack_b200 import _lib
    from pixtrack_b200.refiner import FramePlan

    torch.set_grad_enabled(False)
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lam = lam0().to(dev)
    frames = host_frames(rank)

    # per ring slot: pinned host copies of the per-frame inputs (query pyramid, channels-last),
    # device-resident copies, a device staging area for the end-to-end leg, and prepared launch chains
    host, plans, plans_e2e, staging = [], [], [], []
    for fr in frames:
        h = dict(fq=[f.permute(1, 2, 0).contiguous().pin_memory() for f in fr['F_q']],
                 wq=[w[0].contiguous().pin_memory() for w in fr['W_q']])
        ref = dict(cams=[c.to(dev) for c in fr['cam']], F_ref=[x.to(dev) for x in fr['F_ref']],
                   W_ref=[x.to(dev) for x in fr['W_ref']], p3d=fr['p3d'].to(dev), T_init=fr['T_init'].to(dev))
        fq, wq = [x.to(dev) for x in h['fq']], [x.to(dev) for x in h['wq']]
        st = dict(fq=[torch.empty_like(x) for x in fq], wq=[torch.empty_like(x) for x in wq])
        host.append(h)
        staging.append(st)
        plans.append(FramePlan(fq, wq, lams=[lam] * 3, **ref, **STOP).capture())
        plans_e2e.append(FramePlan(st['fq'], st['wq'], lams=[lam] * 3, **ref, **STOP).capture())
    ring_bytes = sum(x.numel() * 4 for x in host[0]['fq'] + host[0]['wq']) * RING
    graphs = plans[0].graph is not None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K frames, inputs already in HBM ---------------------------------
    for i in range(args.warmup):
        plans[i % RING].run()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        plans[i % RING].run()
    ev1.record()
    barrier()
    n_launch = 3 * args.steps
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    _lib.device_status(local_rank)

    # ---- per-launch timing of the LM kernel for the roofline (events around each bare launch) ------
    per_level = {}
    for i in range(min(args.steps, 4 * RING)):
        pl = plans[i % RING]
        for k, L in enumerate(pl.launches):
            lv = 2 - k
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            L.launch()
            b.record()
            torch.cuda.synchronize()
            Cc = L.shape[2]
            n_it = L.n_iters.cpu()
            lg = L.log[:, :, 1].cpu()
            nv = sum(float(lg[v, :int(n_it[v])].sum()) for v in range(N_VIEWS))
            rec = per_level.setdefault(lv, dict(ms=0.0, bytes=0.0, n=0, iters=0))
            rec['ms'] += a.elapsed_time(b)
            rec['bytes'] += nv * (52 * Cc + 32)
            rec['n'] += 1
            rec['iters'] += int(n_it.max())
    clk = clocks.stop() if rank == 0 else None
    dom = max(per_level, key=lambda k: per_level[k]['ms'])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    ach = per_level[dom]['bytes'] / (per_level[dom]['ms'] * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': f'lm_kernel (pyramid level {dom} launch)', 'achieved': ach, 'peak': peak,
                'unit': 'GB/s', 'frac': ach / peak, 'traffic': None,
                'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650 GB/s',
                'per_level': {str(k): {'us_per_launch': 1e3 * v['ms'] / v['n'], 'GBps': v['bytes'] / (v['ms'] * 1e-3) / 1e9,
                                       'max_iters_per_launch': v['iters'] / v['n']} for k, v in per_level.items()}}
    stress = lm_stress(dev, lam, peak) if rank == 0 else None

    # ---- end to end: host buffers in (pinned -> staging), poses out ---------------------------------
    def e2e_step(i):
        h, st, pl = host[i % RING], staging[i % RING], plans_e2e[i % RING]
        for dst, src in zip(st['fq'] + st['wq'], h['fq'] + h['wq']):
            dst.copy_(src, non_blocking=True)
        pl.run()
        return pl.T.cpu(), pl.failed.cpu()
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

    # ---- final gather of per-frame results (the only collective on this path) -----------------------
    res = torch.stack([torch.cat([pl.T, pl.failed.float()[:, None]], 1) for pl in plans])
    if world > 1:
        gathered = [torch.empty_like(res) for _ in range(world)]
        dist.all_gather(gathered, res)
    iters = [[int(x.max()) for x in pl.n_iters] for pl in plans]
    ok = all(not bool(pl.failed.any()) for pl in plans)

    if rank == 0:
        cpu = None
        if world == 1:
            th = os.cpu_count() or 1
            cpu_refine_view(frames[0], 0, th)
            t0 = time.perf_counter()
            n_v = 0
            while n_v < 40 and time.perf_counter() - t0 < 15:
                cpu_refine_view(frames[n_v // N_VIEWS % RING], n_v % N_VIEWS, th)
                n_v += 1
            dt = time.perf_counter() - t0
            cpu = {'value': n_v / dt / N_VIEWS, 'unit': 'frames/s', 'cores': th, 'kind': 'port',
                   'sample': f'{n_v} view refinements (each 3 levels to convergence) of the same C2 frames via '
                             f'oracle/lm.py in {dt:.1f} s; frames/s = views/s / 8'}
        fps = args.steps * world / (ms_total * 1e-3)
        line = {
            'metric': 'tracked frames/sec (LM-to-convergence)', 'value': fps, 'unit': 'frames/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_total / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD,
                       'l2': f'inputs larger than L2: ring of {RING} frames, {ring_bytes / 1e6:.0f} MB of maps',
                       'stage': 'LM only (pyramids synthetic; extractor and NeRF render not in the timed step yet)',
                       'cuda_graph': graphs, 'lm_iters_coarse_to_fine': iters, 'no_failures': ok},
            'e2e': {'value': args.steps * world / float(e2e_s), 'unit': 'frames/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h},
            'gpu_launches': n_launch, 'clocks': clk, 'roofline': roofline, 'lm_stress': stress, 'cpu_baseline': cpu,
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

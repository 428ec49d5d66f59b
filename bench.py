#!/usr/bin/env python
"""bench.py -- tracked frames/s of the PixTrack pose-refinement hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1] restated on synthetic data, SURVEY 8d "C2"): one step = one
tracked frame of pixloc_tracker_r9.py's refine() minus the NeRF render:
  1. reference refresh: UNet extraction of the re-rendered reference view (1008x756 image) and
     sparse sampling of its 3-level pyramid at the N=5000 model points into one of the B=8 view slots
     (r9.py:154-160 -> extract_reference_features);
  2. query: UNet extraction of the 1920x1080 frame (resized to 1024x576 on the device) and the
     coarse-to-fine LM to convergence (reference stop criteria, num_iters=150) over the 3-level
     pyramid against the B=8 cached reference views (refine_query_pose).
N>1: one process per GPU, independent sequences sharded over ranks (weak scaling), one NCCL
all_gather of the poses at the end.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

sys.path.insert(1, os.path.join(ROOT, 'tests'))     # synthetic scene generators (test / bench infrastructure)
import synthetic as syn  # noqa: E402

N_POINTS, N_VIEWS, RING = 5000, 8, 4
WORKLOAD = ('C2 frame: 1920x1080 query + 1008x756 re-rendered reference view of one textured object; per frame '
            '2 UNet(VGG19) extractions (1024x576 and 1008x756 nets, random-init weights), reference sparse sampling '
            'at N=5000 points into 1 of B=8 view slots, 3-level coarse-to-fine LM to convergence against the 8 views '
            '(num_iters=150, stop 1e-4/5e-3/5e-2, damping const=0); NeRF render excluded (no CPU path in the reference)')
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


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's CPU path (the reference is Python and is not present
# on the GPU box; oracle/ restates it with the same torch CPU primitives: conv2d, grid_sample,
# einsum, linalg.cholesky)
# ------------------------------------------------------------------------------------------------
class CpuFrame:
    def __init__(self, seq, threads):
        from oracle import lm
        torch.set_num_threads(threads)
        self.seq, self.sd, self.lams = seq, syn.unet_weights(0), [lm.damping_lambda(torch.zeros(6))] * 3
        self.obs = {}

    def step(self, i):
        """One frame: reference extraction + sampling, query extraction, B view refinements.  The
        reference keeps B observation sets; only slot i % B is refreshed per frame, the others reuse
        the observations of the frame that filled them (first use fills every slot)."""
        from oracle import lm, unet
        fr = self.seq['frames'][i % len(self.seq['frames'])]
        fr_f, sc_r, cf_r = unet.extract(self.sd, fr['img_r'].numpy().astype(np.float32))
        maps_r = [torch.cat([f, c], 0) for f, c in zip(fr_f, cf_r)]
        obs, keep = lm.sample_reference(maps_r, sc_r, self.seq['cam_r'], fr['R_r'], fr['t_r'], self.seq['p3d'])
        new = ([o[keep] for o in obs], self.seq['p3d'][keep].float())
        for v in range(N_VIEWS):
            if v == i % N_VIEWS or v not in self.obs:
                self.obs[v] = new
        fq, sc_q, cf_q = unet.extract(self.sd, fr['img_q'].numpy().astype(np.float32))
        maps_q = [torch.cat([f, c], 0) for f, c in zip(fq, cf_q)]
        T = []
        for v in range(N_VIEWS):
            T0 = fr['T_init'][v]
            out = lm.refine_levels(maps_q, sc_q, self.seq['cam_q'].float(), T0[:9].reshape(3, 3), T0[9:], self.obs[v][0],
                                   self.obs[v][1], self.lams, **STOP)
            T.append(torch.cat([out['R'].reshape(-1), out['t']]))
        return torch.stack(T)


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_grad_enabled(False)
    threads = os.cpu_count() or 1
    cpu = CpuFrame(syn.tracked_sequence(100, n_frames=RING, N=N_POINTS, n_views=N_VIEWS), threads)
    budget = 240.0
    t_start = time.perf_counter()
    for i in range(min(args.warmup, 1)):
        cpu.step(i)
    t0 = time.perf_counter()
    done = 0
    for i in range(args.steps):
        cpu.step(i)
        done += 1
        if time.perf_counter() - t_start > budget:
            break
    dt = time.perf_counter() - t0
    fps = done / dt
    line = {
        'impl': 'reference', 'metric': 'tracked frames/sec (LM-to-convergence)', 'value': fps, 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': done, 'warmup': min(args.warmup, 1), 'ms_per_step': 1e3 * dt / done,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD},
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{done} full frames (of {args.steps} requested; 240 s budget) through oracle/: 2 UNet '
                                   f'extractions + reference sampling + {N_VIEWS} view refinements each'},
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
    return {'kernel': 'lm_kernel', 'bound': 'hbm',
            'workload': 'C4 level 1: C=128 144x256, N=20000, B=16, 30 fixed iterations', 'ms_per_launch': ms,
            'us_per_iteration': 1e3 * ms / 30, 'achieved': byts / (ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
            'frac': byts / (ms * 1e-3) / 1e9 / peak, 'traffic': lm_dram_traffic(), 'algorithmic_bytes': byts,
            'ctas_per_problem': g, 'problems_in_flight': ng,
            'note': 'algorithmic bytes (52C+32 per valid point per iteration); the 19 MB map stays L2-resident, so '
                    'this is L2-served traffic measured against the HBM copy peak'}


def lm_dram_traffic():
    """DRAM bytes of one lm_kernel launch of the C4 stress problem from the committed `ncu --set full` capture."""
    try:
        r = json.load(open(os.path.join(ROOT, 'profiles', 'r1_lm_v3_ncu.json')))['launches'][0]
    except (OSError, KeyError, ValueError, IndexError):
        return None
    mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tot = 0.0
    for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        v, u = r[k].split()
        tot += float(v) * mult[u]
    return tot


def conv_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the plan's tensor-core conv launches, summed, from the committed
    `ncu --set full` capture of profiles/extractor_profile.py (cold caches, so an upper bound; None when absent)."""
    path = os.path.join(ROOT, 'profiles', 'r1_conv_final_ncu.json')
    try:
        rows = json.load(open(path))['launches']
    except (OSError, KeyError, ValueError):
        return None
    mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tot = 0.0
    for r in rows:
        for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            v, u = r[k].split()
            tot += float(v) * mult[u]
    return {'bytes': tot, 'launches': len(rows), 'source': 'profiles/r1_conv_final_ncu.json'}


def nerf_leg(dev, frame_step, steps):
    """Reference-view re-render (1008x756, spp 8: r9.py:81,150 + run_vis_on_poses.py:29) of a random-weight
    instant-ngp model with a ball-shaped occupancy, alone and in front of the tracked frame."""
    from pixtrack_b200.nerf import NerfTestbed, occupancy_bitfield
    sc = syn.nerf_scene(11, 2)
    tb = NerfTestbed(sc['grid'], sc['w_density'], sc['w_rgb'], occupancy_bitfield(sc['density_grid'], sc['max_cascade']),
                     2, dev)
    tb.nerf.rendering_min_transmittance = 1e-7
    tb.fov = 40.0
    tb.set_ngp_camera_matrix(syn.nerf_look_at((0.4, -1.3, 0.8)))
    W, H, spp = 1008, 756, 8
    for _ in range(3):
        rgba, u8, _ = tb.render_device(W, H, spp, want_u8=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    a.record()
    for _ in range(reps):
        tb.render_device(W, H, spp, want_rgba=False, want_u8=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    a.record()
    for i in range(steps):
        tb.render_device(W, H, spp, want_rgba=False, want_u8=True)
        frame_step(i)
    b.record()
    torch.cuda.synchronize()
    ms_frame = a.elapsed_time(b) / steps
    cover = float((rgba[..., 3] > 0.5).float().mean())
    return {'workload': f'{W}x{H} spp {spp}, aabb_scale 2, ball occupancy covering {cover:.0%} of the image, random weights',
            'ms_per_render': ms, 'mrays_per_s': W * H * spp / ms / 1e3, 'frames_per_s_with_render': 1e3 / ms_frame,
            'ms_per_frame_with_render': ms_frame}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from pixtrack_b200 import _lib, shard
    from pixtrack_b200.extractor import B200FeatureExtractor
    from pixtrack_b200.pipeline import FrameTracker

    torch.set_grad_enabled(False)
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner with printf while the communicator is
        # created, so file descriptor 1 points at stderr during the (eager) initialisation and a first collective
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    lam = lam0().to(dev)
    seq = syn.tracked_sequence(100 + rank, n_frames=RING, N=N_POINTS, n_views=N_VIEWS)
    frames = seq['frames']
    ext = B200FeatureExtractor(syn.unet_weights(0), dev)
    trk = FrameTracker(ext, frames[0]['img_q'].shape[:2], seq['cam_q'], seq['p3d'], [lam] * 3, N_VIEWS, **STOP)

    host = [dict(q=f['img_q'].pin_memory(), r=f['img_r'].pin_memory()) for f in frames]
    devi = [dict(q=h['q'].to(dev), r=h['r'].to(dev)) for h in host]
    stage = dict(q=torch.empty_like(devi[0]['q']), r=torch.empty_like(devi[0]['r']))
    T_ref = [torch.cat([f['R_r'].reshape(-1), f['t_r']]) for f in frames]
    T_init = [f['T_init'].to(dev) for f in frames]
    for v in range(N_VIEWS):            # fill every view slot once (untimed set-up)
        trk.refresh_reference(v, devi[0]['r'], seq['cam_r'], T_ref[0])

    def step(i, imgs):
        k = i % RING
        trk.refresh_reference(i % N_VIEWS, imgs['r'], seq['cam_r'], T_ref[k])
        return trk.track(imgs['q'], T_init[k])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K frames, images already in HBM ---------------------------------
    for i in range(args.warmup):
        step(i, devi[i % RING])
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(i, devi[i % RING])
    ev1.record()
    barrier()
    ms_total = shard.max_over_ranks(ev0.elapsed_time(ev1), dev)
    _lib.device_status(local_rank)
    clk = clocks.stop() if rank == 0 else None

    # ---- results of the last ring pass: LM iteration counts, failures, pose error vs ground truth ------
    iters, errs, ok = [], [], True
    for k in range(RING):
        T, failed = step(k, devi[k])
        n_it = [int(x.max()) for x in trk.plan.n_iters]
        torch.cuda.synchronize()
        iters.append(n_it)
        ok = ok and not bool(failed.any())
        Tc = T.double().cpu()
        dR = Tc[:, :9].reshape(-1, 3, 3) @ frames[k]['R_q'].t()
        ang = torch.rad2deg(torch.acos(((dR.diagonal(dim1=1, dim2=2).sum(-1) - 1) / 2).clamp(-1, 1)))
        errs.append([float(ang.median()), float((Tc[:, 9:] - frames[k]['t_q']).norm(dim=1).median())])

    # ---- rooflines (rank 0): per-launch CUDA-event timing of the extractor plan; LM stress ---------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    roofline = stress = plan_prof = None
    if rank == 0:
        rows = None
        for _ in range(3):
            r = ext.profile(devi[0]['q'])
            rows = r if rows is None else [(a[0], min(a[1], b[1]), a[2]) for a, b in zip(rows, r)]
        tc_ms = sum(r[1] for r in rows if r[0] == 'conv_tc')
        tc_fl = sum(r[2] for r in rows if r[0] == 'conv_tc')
        n_tc = sum(1 for r in rows if r[0] == 'conv_tc')
        peak_tf = float(peaks.get('bf16_tflops', 1590.0))
        ach = tc_fl / (tc_ms * 1e-3) / 1e12
        by_kind = {}
        for k, m, f in rows:
            d = by_kind.setdefault(k, [0.0, 0.0, 0])
            d[0] += m
            d[1] += f
            d[2] += 1
        plan_prof = {k: {'launches': v[2], 'ms': v[0], 'gflop': v[1] / 1e9} for k, v in by_kind.items()}
        roofline = {'bound': 'tensor', 'kernel': f'conv_halo_kernel / conv_tc_kernel (tcgen05 implicit-GEMM convs; the {n_tc} '
                    'launches of the 1024x576 plan = 557 of its 561 GFLOP)', 'achieved': ach, 'peak': peak_tf,
                    'unit': 'TFLOP/s', 'frac': ach / peak_tf, 'traffic': conv_dram_traffic(),
                    'peak_source': ('MEASURED_PEAKS.json bf16_tflops (burst; kernels timed one by one with CUDA events)'
                                    if peaks else 'fallback 1590 TFLOP/s'),
                    'avg_launch_us': 1e3 * tc_ms / n_tc, 'share_of_plan_time': tc_ms / sum(r[1] for r in rows)}
        stress = lm_stress(dev, lam, float(peaks.get('hbm_gbs', 6650.0)))

    # ---- end to end: host images in (pinned -> device staging), poses out --------------------------------
    # Camera frames do not depend on the pose, so the upload of frame i+1 runs on a copy stream while frame i is
    # tracked (two staging sets); every step still uploads its own inputs inside the timed region and ends with a
    # device->host read of its poses (the tracker needs them on the host for the next render / reference choice).
    stages = [stage, dict(q=torch.empty_like(stage['q']), r=torch.empty_like(stage['r']))]
    copy_stream = torch.cuda.Stream(dev)
    uploaded = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(i):
        st, h = stages[i & 1], host[i % RING]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])      # frame i-2 no longer reads this staging set
            st['q'].copy_(h['q'], non_blocking=True)
            st['r'].copy_(h['r'], non_blocking=True)
            uploaded[i & 1].record(copy_stream)

    def e2e_run(n):
        for ev in consumed:
            ev.record(torch.cuda.current_stream(dev))
        upload(0)
        res = torch.empty(N_VIEWS * 13, dtype=torch.float32).pin_memory()
        for i in range(n):
            torch.cuda.current_stream(dev).wait_event(uploaded[i & 1])
            T, failed = step(i, stages[i & 1])
            consumed[i & 1].record(torch.cuda.current_stream(dev))
            if i + 1 < n:
                upload(i + 1)                # enqueued behind this frame's launches; copies while it computes
            res[:N_VIEWS * 12].copy_(T.reshape(-1), non_blocking=True)
            res[N_VIEWS * 12:].copy_(failed.reshape(-1), non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()     # the poses are on the host before the next frame starts
    e2e_run(max(2, args.warmup // 2))
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_s = shard.max_over_ranks(time.perf_counter() - t0, dev)

    # ---- the same frame with the NeRF re-render of the reference view in front (r9.py:145-160 renders it
    #      every frame; reported separately because the reference has no CPU path for it) ---------------
    nerf = nerf_leg(dev, lambda i: step(i, devi[i % RING]), args.steps) if rank == 0 else None
    h2d = host[0]['q'].numel() + host[0]['r'].numel()
    d2h = N_VIEWS * 13 * 4           # [B,12] fp32 poses + B failure flags, read back as one pinned fp32 buffer

    # ---- final gather of per-unit results (the only collective on this path): unit = this rank's sequence,
    #      its result = the pose of the best view --------------------------------------------------------
    best = 0
    table = shard.gather_results(shard.pack_results([rank], trk.plan.T[best:best + 1], trk.plan.failed[best:best + 1],
                                                    trk.plan.n_iters[-1][best:best + 1]), world)
    assert table.shape[0] == world

    if rank == 0:
        cpu = None
        if world == 1:
            th = os.cpu_count() or 1
            cf = CpuFrame(seq, th)
            t0 = time.perf_counter()
            n_f = 0
            while n_f < 8 and (n_f == 0 or time.perf_counter() - t0 < 12):     # bounded sample: about 12-15 s of CPU work
                cf.step(n_f)
                n_f += 1
            dt = time.perf_counter() - t0
            cpu = {'value': n_f / dt, 'unit': 'frames/s', 'cores': th, 'kind': 'port',
                   'sample': f'{n_f} full frame(s) of the same sequence through oracle/ (2 UNet extractions + reference '
                             f'sampling + {N_VIEWS} view refinements each) in {dt:.1f} s'}
        fps = args.steps * world / (ms_total * 1e-3)
        line = {
            'metric': 'tracked frames/sec (LM-to-convergence)', 'value': fps, 'unit': 'frames/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_total / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16+f32',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD,
                       'arithmetic': 'extractor: fp16 tensor-core operands, fp32 accumulation; sampler projection f64; LM f32',
                       'l2': f'inputs larger than L2: ring of {RING} distinct frames; one frame streams >600 MB of '
                             'activations and maps through the 126 MB L2',
                       'cuda_graph': trk.plan.graph is not None, 'lm_iters_coarse_to_fine': iters, 'no_failures': ok,
                       'median_pose_error_deg_m_vs_gt': errs},
            'e2e': {'value': args.steps * world / float(e2e_s), 'unit': 'frames/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h},
            # launches per frame: 2 extractor plans, 1 sampler, 3 LM launches (one graph)
            'gpu_launches': (2 * len(rows) + 1 + 3) * args.steps, 'clocks': clk, 'roofline': roofline,
            'roofline_lm': stress, 'extractor_plan': plan_prof, 'nerf_render': nerf, 'cpu_baseline': cpu,
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
    ap.add_argument('--workload', default='c2', choices=['c2', 'c4'],
                    help='c2 (default, the configuration the metric is quoted on) or c4: BASELINE.json configs[3] '
                         '(16 views, 20000 points, 30 fixed LM iterations per level, frames sharded over the GPUs)')
    args = ap.parse_args()
    if args.workload == 'c4':
        global N_POINTS, N_VIEWS, STOP, WORKLOAD
        N_POINTS, N_VIEWS = 20000, 16
        STOP = dict(num_iters=30, grad_stop=0.0, dt_stop=0.0, dR_stop=0.0)
        WORKLOAD = ('C4 frame: synthetic 1920x1080 query + 1008x756 reference view; per frame 2 UNet(VGG19) extractions, '
                    'reference sparse sampling at N=20000 points into 1 of B=16 view slots, 3-level LM with 30 fixed '
                    'iterations per level against the 16 views; independent frames sharded over the GPUs')
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

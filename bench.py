#!/usr/bin/env python
"""bench.py -- frames/s of the SRFDet3D point-cloud -> region-feature hot path on B200.

  python bench.py --gpus N --steps K --warmup W [--workload nusc_L|nusc_LC|waymo_L|kitti_L]
  torchrun ... bench.py --gpus N ...          (one rank per GPU; frames are independent)
  python bench.py --impl reference ...        (restated reference CPU path on the host cores)

A step = ONE frame per rank: hard/dynamic voxelization (+VFE) -> SparseEncoder -> 5 stages
of region fusion (BEV RoIAlign [+ 6-camera image RoIAlign + fusion Linear] -> DynamicConv).
`value` = frames/s over all ranks with the point clouds already resident in HBM;
`e2e` = the same through the public host-buffer call (pinned host points in, host region
features out, H2D + D2H inside the timed region).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    'nusc_L': dict(kind='nusc', fusion=False, desc='srfdet_voxel_nusc_L: 300k-pt 10-sweep nuScenes-shaped cloud, hard voxelization '
                   '0.075/0.075/0.2 m (1472x1472x41), 21-layer SparseEncoder, 900 proposals x 5 stages BEV RoIAlign + DynamicConv'),
    'nusc_LC': dict(kind='nusc', fusion=True, desc='srfdet_voxel_nusc_LC: nusc_L + 6-view image RoIAlign (1600x928) + fusion Linear'),
    'waymo_L': dict(kind='waymo', fusion=False, desc='srfdet_dvoxel_waymo_L: 180k-pt cloud, dynamic voxelization + DynamicVFE'),
    'kitti_L': dict(kind='kitti', fusion=False, desc='srfdet_voxel_kitti_L: 120k-pt cloud, dynamic voxelization + DynamicVFE, C=256 head'),
}
N_CLOUDS = 8


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sust=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


class ClockSampler:
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = f'/tmp/srf_clocks_{os.getpid()}.csv'

    def start(self):
        try:
            self.f = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                                          '-i', str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))


def conv_roofline(pipe, pts, pk, torch):
    """Per-launch CUDA-event timing of every sparse-conv launch of one frame (events on the
    launching stream), with the ALGORITHMIC flops / bytes of each layer from its rulebook."""
    enc = pipe.detector.pts_middle_encoder
    for _ in range(2):
        pipe.encode(pts)
    rows = None
    reps = 5
    for r in range(reps):
        enc.profile = []
        # a spin kernel keeps the GPU busy while the host enqueues the whole encoder, so the events
        # bracket kernel execution only (no host launch latency between an event and its kernel)
        torch.cuda._sleep(30_000_000)
        pipe.encode(pts)
        torch.cuda.synchronize()
        prof, enc.profile = enc.profile, None
        if rows is None:
            rows = []
            for e in prof:
                n_out, n_in = int(e['n_out']), int(e['n_in'])
                pairs = int((e['nbr'][:, :n_out] >= 0).sum())
                flops = 2.0 * pairs * e['cin'] * e['cout']
                byts = n_in * e['cin'] * e['in_bytes'] + n_out * e['cout'] * e['out_bytes'] + \
                    e['kvol'] * e['cin'] * e['cout'] * e['in_bytes'] + 4.0 * e['kvol'] * n_out
                rows.append(dict(layer=e['layer'], kernel='igemm_umma' if e['umma'] else 'spconv_f32', subm=e['subm'],
                                 cin=e['cin'], cout=e['cout'], kvol=e['kvol'], n_in=n_in, n_out=n_out, pairs=pairs,
                                 flops=flops, bytes=byts, ms=[]))
        for row, e in zip(rows, prof):
            row['ms'].append(e['start'].elapsed_time(e['end']))
    for row in rows:
        ms = sorted(row.pop('ms'))
        row['ms'] = ms[len(ms) // 2]
        row['tflops'] = row['flops'] / (row['ms'] * 1e-3) / 1e12
        row['gbs'] = row['bytes'] / (row['ms'] * 1e-3) / 1e9
    umma = [r for r in rows if r['kernel'] == 'igemm_umma']
    tot_ms = sum(r['ms'] for r in umma)
    tot_fl = sum(r['flops'] for r in umma)
    # dominant kernel = the instantiation with the largest share of the step; its numbers are per-launch averages
    groups = {}
    for r in (umma or rows):
        groups.setdefault((r['cin'], r['cout']), []).append(r)
    key, grp = max(groups.items(), key=lambda kv: sum(r['ms'] for r in kv[1]))
    n = len(grp)
    ms = sum(r['ms'] for r in grp) / n
    flops = sum(r['flops'] for r in grp) / n
    byts = sum(r['bytes'] for r in grp) / n
    tflops, gbs = flops / (ms * 1e-3) / 1e12, byts / (ms * 1e-3) / 1e9
    # roofline side: arithmetic intensity against the ridge of the two measured peaks
    ridge = pk['tf_sust'] * 1e12 / (pk['hbm'] * 1e9)
    hbm_bound = flops / byts < ridge
    name = f"igemm_umma_kernel<{key[0]},{key[1]}>"
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'r01_ncu_traffic.json')) as f:
            t = json.load(f)['kernels'].get(name)
        if t:
            traffic = t['dram_read_bytes'] + t['dram_write_bytes']
            traffic_src = 'profiles/r01_ncu_traffic.json (ncu --set full, dram read+write per launch, cold cache)'
    except (OSError, ValueError, KeyError):
        pass
    roof = dict(bound='hbm' if hbm_bound else 'tensor',
                kernel=f"{name} ({n} launches per frame: layers {[r['layer'] for r in grp]}, SubM k27, "
                       f"{grp[0]['n_out']} rows, {grp[0]['pairs']} pairs)",
                achieved=round(gbs if hbm_bound else tflops, 2), peak=pk['hbm'] if hbm_bound else pk['tf_sust'],
                unit='GB/s' if hbm_bound else 'TFLOP/s',
                frac=round(gbs / pk['hbm'] if hbm_bound else tflops / pk['tf_sust'], 4),
                traffic=traffic, traffic_source=traffic_src,
                peak_source=f"{pk['src']} {'HBM copy bandwidth' if hbm_bound else 'bf16 sustained'}", launch_ms=round(ms, 4),
                algorithmic_flops_per_launch=flops, algorithmic_bytes_per_launch=byts,
                arithmetic_intensity=round(flops / byts, 1), ridge=round(ridge, 1),
                tensor_frac_of_same_launch=round(tflops / pk['tf_sust'], 4), hbm_frac_of_same_launch=round(gbs / pk['hbm'], 4),
                share_of_step_ms=round(sum(r['ms'] for r in grp), 4),
                all_umma_launches=dict(n=len(umma), ms=round(tot_ms, 4), tflops=round(tot_fl / (tot_ms * 1e-3) / 1e12, 2) if tot_ms else None))
    return roof, rows


def cpu_frame_seconds(state, kind, d, pts_np, torch):
    from oracle import cpu_pipeline
    from srfdet_b200 import synth
    t0 = time.perf_counter()
    cpu_pipeline.run_frame(state, kind, synth.GEOM[kind], d, pts_np)
    return time.perf_counter() - t0


def run_reference(args, wl):
    """--impl reference: the restated reference CPU path (oracle) on the host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from oracle import oracle as O
    O.build_c()
    from srfdet_b200 import synth
    from srfdet_b200 import pipeline as P
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind = wl['kind']
    state = build_cpu_state(kind, wl['fusion'], torch)
    d = P.HEAD_CFG[kind]['d']
    clouds = [synth.cloud(kind, 1000 + i) for i in range(2)]
    for i in range(args.warmup):
        cpu_frame_seconds(state, kind, d, clouds[i % 2], torch)
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_frame_seconds(state, kind, d, clouds[i % 2], torch)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    line = dict(metric='frames/s (voxelize+SparseEncoder+RoI fusion)', value=round(fps, 4), unit='frames/s', n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=round(dt / args.steps * 1e3, 2), higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=dict(workload=args.workload, description=wl['desc'], frames_per_step=1),
                cpu_baseline=dict(value=round(fps, 4), unit='frames/s', cores=cores, kind='port',
                                  sample=f'{args.steps} full frames ({clouds[0].shape[0]} points each) through the oracle port '
                                         '(mmcv CPU voxelize loop in C, spconv native gather-mm-scatter with torch.mm, RoIAlign C loop, torch CPU bmm)'),
                e2e=dict(value=round(fps, 4), unit='frames/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def build_cpu_state(kind, fusion, torch):
    """Same seeded weights / synthetic maps as RegionFeaturePipeline, built without CUDA."""
    from srfdet_b200.pipeline import RegionFeaturePipeline
    pipe = RegionFeaturePipeline(kind, fusion=fusion, device='cpu', precision='fp32')
    return pipe.state()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='nusc_L', choices=sorted(WORKLOADS))
    ap.add_argument('--precision', default='fp16', choices=['fp16', 'fp32', 'bf16', 'fp32_simt'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel eagerly instead of replaying a CUDA graph of the frame')
    ap.add_argument('--nchw', action='store_true', help='RoI stage samples contiguous NCHW maps (reference layout) instead of channels_last')
    ap.add_argument('--layers-out', default=None, help='write the per-layer sparse-conv table (JSON) here')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    from srfdet_b200 import _lib as L
    from srfdet_b200 import frames, synth
    from srfdet_b200.pipeline import RegionFeaturePipeline
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU fallback)'
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    lib = L.load()
    pk = peaks()
    kind = wl['kind']
    pipe = RegionFeaturePipeline(kind, fusion=wl['fusion'], precision=args.precision, channels_last=not args.nchw)
    eager = pipe._run_frame_eager
    # distinct frames per rank, resident on the device (value) and in pinned host memory (e2e)
    clouds_np = [synth.cloud(kind, 1000 * (rank + 1) + i) for i in range(N_CLOUDS)]
    clouds_dev = [torch.as_tensor(c).cuda() for c in clouds_np]
    clouds_pin = [torch.as_tensor(c).pin_memory() for c in clouds_np]
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device='cuda')   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, inputs, steps, warmup):
        for i in range(warmup):
            fn(inputs[i % N_CLOUDS])
        barrier()
        evs = []
        for i in range(steps):
            flush.zero_()                                # evict the previous frame from L2 (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(inputs[i % N_CLOUDS])
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        return frames.max_over_ranks(ms, 'cuda')

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eager(clouds_dev[0])
    n0 = lib.srf_launch_count()
    eager(clouds_dev[0])
    launches_per_frame = int(lib.srf_launch_count() - n0)      # this library's kernels per frame (counted eagerly)
    pipe.use_graph = not args.no_graph
    ms_dev = timed(pipe.run_frame, clouds_dev, args.steps, args.warmup)
    ms_e2e = timed(pipe.run_frame_host, clouds_pin, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    fps = world * args.steps / (ms_dev * 1e-3)
    fps_e2e = world * args.steps / (ms_e2e * 1e-3)

    roof = rows = cpu = None
    pipe.use_graph = False
    if rank == 0:
        roof, rows = conv_roofline(pipe, clouds_dev[0], pk, torch)
        if args.layers_out:
            with open(args.layers_out, 'w') as f:
                json.dump(dict(workload=args.workload, precision=args.precision, layers=rows), f, indent=1)
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            state = pipe.state()
            cpu_frame_seconds(state, kind, pipe.d, synth.cloud(kind, 1, n_points=20000), torch)   # warm caches / build
            t = cpu_frame_seconds(state, kind, pipe.d, clouds_np[0], torch)
            cpu = dict(value=round(1.0 / t, 4), unit='frames/s', cores=cores, kind='port',
                       sample=f'1 full frame ({clouds_np[0].shape[0]} points) through the oracle port of the reference CPU path, {t:.1f} s')
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    n_pts, c_pts = clouds_np[0].shape
    line = dict(metric='frames/s (voxelize+SparseEncoder+RoI fusion)', value=round(fps, 2), unit='frames/s', n_gpus=world,
                steps=args.steps, warmup=args.warmup, ms_per_step=round(ms_dev / args.steps, 4), higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype={'fp16': 'f16', 'bf16': 'bf16', 'fp32': 'f16x2 (hi+lo split operands, 3 MMAs per product, fp32 accumulate)', 'fp32_simt': 'f32'}[args.precision], data='synthetic',
                config=dict(workload=args.workload, description=wl['desc'], frames_per_step_per_gpu=1,
                            parallelism=f'{world} independent frame replica(s), no data-path collective',
                            l2='512 MiB flush written between timed steps; 8 distinct clouds cycled',
                            roi_map_layout='NCHW contiguous' if args.nchw else 'torch.channels_last (NHWC in memory)',
                            launch='eager' if args.no_graph else 'one CUDA graph per frame (captured once, replayed)',
                            excluded='dense BEV backbone/FPN, image backbone, attention/FFN rows of the head (SURVEY 8f): RoI stage samples synthetic FPN maps'),
                clocks=clocks,
                e2e=dict(value=round(fps_e2e, 2), unit='frames/s', h2d_bytes_per_step=int(n_pts * c_pts * 4),
                         d2h_bytes_per_step=int(900 * pipe.C * 4), ms_per_step=round(ms_e2e / args.steps, 4)),
                gpu_launches=launches_per_frame * args.steps,
                gpu_launches_per_step=launches_per_frame,
                roofline=roof, cpu_baseline=cpu)
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()

"""One eagerly launched frame for `ncu --profile-from-start off`: the profiler range is opened only around the
first `--per-key` occurrences of every distinct C-ABI call signature (entry point + shape summary), so a
`--set full` capture costs ~50 kernels instead of every launch of the frame.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/frame \
        python tools/ncu_frame.py [--workload nusc_LC] [--precision fp16] [--only REGEX] [--per-key 1]
    python tools/ncu_frame.py --dry          # list the selected calls without a profiler

Without ncu attached cudaProfilerStart/Stop are no-ops, so the same command is the "exits 0 without ncu" check.
"""
import argparse
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='nusc_LC')
    ap.add_argument('--precision', default='fp16')
    ap.add_argument('--only', default=None, help='regex over "entry family": profile only matching calls')
    ap.add_argument('--per-key', type=int, default=1)
    ap.add_argument('--dry', action='store_true')
    args = ap.parse_args()
    import torch
    import bench
    from srfdet_b200 import _lib as L
    from srfdet_b200 import profiling, synth
    from srfdet_b200.pipeline import RegionFeaturePipeline
    w = bench.WORKLOADS[args.workload]
    pipe = RegionFeaturePipeline(w['kind'], fusion=w['fusion'], precision=args.precision, scope='full')
    pipe.calibrate(torch.as_tensor(synth.cloud(w['kind'], 999)).cuda())
    pts = torch.as_tensor(synth.cloud(w['kind'], 1000)).cuda()
    enc = pipe.detector.pts_middle_encoder
    enc.overlap_geometry = pipe.overlap_stage = pipe.overlap_image_branch = False     # one kernel at a time
    pipe.use_graph = False
    for _ in range(2):
        pipe._run_frame_eager(pts)
    torch.cuda.synchronize()
    lib = L.load()
    rt = torch.cuda.cudart()
    pat = re.compile(args.only) if args.only else None
    seen = {}
    picked = []
    originals = {}
    skip = ('_bytes', '_splits', '_splits_enc', '_tile_k', '_tile_n', '_tile_k_enc')
    for name in L.PROTOTYPES:
        fn = getattr(lib, name)
        originals[name] = fn
        if name.endswith(skip) or name in ('srf_version', 'srf_last_error', 'srf_sm_count', 'srf_launch_count', 'srf_geom_init', 'srf_conv3x3_last_used_tma'):
            continue

        def wrapped(*a, _fn=fn, _name=name):
            ints = tuple(profiling._iv(x) for x in a)
            a0 = getattr(a[0], '_obj', None) if a else None
            summ = None
            if isinstance(a0, L.ConvArgs):
                summ = dict(cin=a0.cin, cout=a0.cout, kvol=a0.kvol, cap=a0.cap_out, in_enc=a0.in_dtype, out_enc=a0.out_dtype,
                            dense=bool(a0.dense), in_rows=a0.in_rows)
            elif isinstance(a0, L.LinearArgs):
                summ = dict(enc=a0.a_enc, m=a0.m, k=a0.k, n=a0.n, out_enc=a0.out_enc, out2=bool(a0.out2), out2_enc=a0.out2_enc,
                            k_splits=max(1, a0.k_splits))
            elif isinstance(a0, L.Pyramid):
                summ = dict(channels=a0.channels, levels=a0.n_levels, hw=[(a0.h[i], a0.w[i]) for i in range(a0.n_levels)])
            elif isinstance(a0, L.Map):
                summ = dict(ca=a0.c, cb=0)
            try:
                fam = profiling._family(_name, ints, summ)
            except Exception:
                fam = 'other'
            key = (_name, fam, repr(summ) if summ else repr(tuple(x for x in ints if isinstance(x, (int, tuple)) and (isinstance(x, tuple) or abs(x) < 1 << 24))))
            n = seen.get(key, 0)
            seen[key] = n + 1
            take = n < args.per_key and fam is not None and (pat is None or pat.search(f'{_name} {fam}'))
            if take:
                picked.append(f'{_name:28s} {fam}')
                if not args.dry:
                    torch.cuda.synchronize()
                    rt.cudaProfilerStart()
            rc = _fn(*a)
            if take and not args.dry:
                torch.cuda.synchronize()
                rt.cudaProfilerStop()
            return rc
        setattr(lib, name, wrapped)
    pipe._run_frame_eager(pts)
    torch.cuda.synchronize()
    for name, fn in originals.items():
        setattr(lib, name, fn)
    print(f'{len(picked)} calls inside the profiler range:')
    for p in picked:
        print('  ', p)


if __name__ == '__main__':
    main()

"""Full-size parity of every precision mode against the CPU oracle (max |a-b| / max |b|).

    python tools/precision_report.py [--workloads nusc_L,nusc_LC,waymo_L,kitti_L] [--points N] [--out FILE]

Prints / writes one JSON object: {workload: {precision: {bev, obj, bev_rms}}}."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cpu_pipeline  # noqa: E402
from oracle import oracle as O  # noqa: E402
from srfdet_b200 import synth  # noqa: E402
from srfdet_b200.pipeline import RegionFeaturePipeline  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from util import box_errors, teacher_forced  # noqa: E402

WL = {'nusc_L': ('nusc', False), 'nusc_LC': ('nusc', True), 'waymo_L': ('waymo', False), 'kitti_L': ('kitti', False)}


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workloads', default='nusc_L,nusc_LC,waymo_L,kitti_L')
    ap.add_argument('--precisions', default='fp32,fp16,bf16,fp32_simt')
    ap.add_argument('--points', type=int, default=0, help='0 = the configuration\'s full cloud')
    ap.add_argument('--scope', default='full', choices=['full', 'path'])
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    O.build_c()
    torch.set_num_threads(os.cpu_count() or 1)
    res = {}
    for wl in args.workloads.split(','):
        kind, fusion = WL[wl]
        pipe = RegionFeaturePipeline(kind, fusion=fusion, precision='fp32', scope=args.scope)
        pts = synth.cloud(kind, 1000, n_points=args.points or None)
        if args.scope == 'full':
            pipe.calibrate(torch.as_tensor(synth.cloud(kind, 999, n_points=args.points or None)).cuda())
        t0 = time.perf_counter()
        state = pipe.state()
        ref_bev = cpu_pipeline.encode(state, kind, synth.GEOM[kind], pts)
        if args.scope == 'full':
            ref_obj, ref_x = cpu_pipeline.full_chain(state, ref_bev)
        else:
            ref_obj, ref_x = cpu_pipeline.region_stages(state, synth.GEOM[kind], pipe.d), None
        t_cpu = time.perf_counter() - t0
        res[wl] = dict(points=int(pts.shape[0]), oracle_s=round(t_cpu, 2))
        for prec in args.precisions.split(','):
            pipe.precision = prec
            bev, obj = pipe.run_frame(torch.as_tensor(pts).cuda())
            torch.cuda.synchronize()
            b, o = bev.cpu().numpy(), obj.cpu().numpy()
            res[wl][prec] = dict(bev=rel(b, ref_bev), obj=rel(o, ref_obj),
                                 bev_rms=float(np.sqrt(((b - ref_bev) ** 2).mean()) / np.sqrt((ref_bev ** 2).mean())))
            if ref_x is not None:
                res[wl][prec].update(box_errors(pipe.last, ref_x, synth.GEOM[kind]['pc_range']))
                res[wl][prec].update(teacher_forced(pipe, ref_x, prec))
            print(wl, prec, res[wl][prec], flush=True)
    s = json.dumps(res, indent=1)
    if args.out:
        open(args.out, 'w').write(s)
    print(s)


if __name__ == '__main__':
    main()

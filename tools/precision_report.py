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

WL = {'nusc_L': ('nusc', False), 'nusc_LC': ('nusc', True), 'waymo_L': ('waymo', False), 'kitti_L': ('kitti', False)}


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def box_errors(last, ref, pc):
    """Errors of the chained head outputs: FPN level 0, logits of all stages, boxes of all stages split into
    centres (as a fraction of the range), log sizes and (sin, cos, v) -- absolute differences."""
    span = np.array([pc[3] - pc[0], pc[4] - pc[1], pc[5] - pc[2]], np.float32)
    gb, rb = last['boxes'].cpu().numpy(), ref['boxes']
    stages = last['logits'].shape[0]
    lg, rl = last['logits'].cpu().numpy(), ref['logits']
    return dict(logits_per_stage=[rel(lg[i], rl[i]) for i in range(stages)], fpn0=rel(last['pyramid'][0].cpu().numpy(), ref['pyramid'][0]), fpn3=rel(last['pyramid'][3].cpu().numpy(), ref['pyramid'][3]),
                logits=rel(last['logits'].cpu().numpy(), ref['logits']),
                centre_frac=float((np.abs(gb[..., :3] - rb[..., :3]) / span).max()), logsize_abs=float(np.abs(gb[..., 3:6] - rb[..., 3:6]).max()),
                rest_abs=float(np.abs(gb[..., 6:] - rb[..., 6:]).max()))


def teacher_forced(pipe, ref, prec):
    """Every stage of the head fed with the ORACLE's inputs of that stage (FPN pyramid, boxes, proposal features):
    isolates each stage's own error from the amplification of the chained loop."""
    tr = ref['trace']
    head = pipe.head
    pyr = [torch.as_tensor(p).cuda().contiguous(memory_format=torch.channels_last) for p in ref['pyramid']]
    img = pipe.img_feats if pipe.fusion else None
    b0, f0 = head._get_init_proposals(img, pyr, sigmoid_centres=True)
    out = dict(tf_init_boxes=float(np.abs(b0.cpu().numpy() - tr['init_boxes']).max()), tf_init_prop=rel(f0[0].cpu().numpy(), tr['init_prop']),
               tf_logits=[], tf_boxes=[], tf_obj=[])
    for s, stage in enumerate(head.head_series_lidar):
        boxes = torch.as_tensor(tr['stage_in'][s][0]).cuda().contiguous()
        prop = torch.as_tensor(tr['stage_in'][s][1]).cuda().contiguous()
        if pipe.fusion:
            lg, pred, obj = stage(img, pyr, boxes, prop, head.roi_extractor_lidar, None, pooler_img=head.roi_extractor_img, precision=prec,
                                  lidar2img=pipe.lidar2img)
        else:
            lg, pred, obj = stage(pyr, boxes, prop, head.roi_extractor_lidar, None, precision=prec)
        rl, rp, ro = tr['stage_out'][s]
        out['tf_logits'].append(rel(lg[0].cpu().numpy(), rl))
        out['tf_boxes'].append(float(np.abs(pred[0].cpu().numpy() - rp).max()))
        out['tf_obj'].append(rel(obj.cpu().numpy(), ro))
    return out


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

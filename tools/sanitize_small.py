"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck)."""
import sys
sys.path.insert(0, '.')
import torch
from srfdet_b200 import synth
from srfdet_b200.pipeline import RegionFeaturePipeline
for kind, fusion in [('nusc', True), ('waymo', False)]:
    for prec in ['bf16', 'fp32']:
        pipe = RegionFeaturePipeline(kind, fusion=fusion, precision=prec)
        pts = torch.as_tensor(synth.cloud(kind, 7, n_points=6000)).cuda()
        bev, obj = pipe.run_frame(pts)
        torch.cuda.synchronize()
        print(kind, prec, float(bev.abs().sum()), float(obj.abs().sum()))
print('sanitize run done')

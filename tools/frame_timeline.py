"""Kernel timeline of one eager frame (torch.profiler / CUPTI): start, duration, gap to the previous kernel
on the same stream.  Development tool."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srfdet_b200 import synth  # noqa: E402
from srfdet_b200.pipeline import RegionFeaturePipeline  # noqa: E402


def main():
    fusion = 'LC' in sys.argv
    pipe = RegionFeaturePipeline('nusc', fusion=fusion, precision='fp16')
    pts = torch.as_tensor(synth.cloud('nusc', 1)).cuda()
    for _ in range(3):
        pipe.run_frame(pts)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        torch.cuda._sleep(40_000_000)
        pipe.run_frame(pts)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    evs = [e for e in evs if 'sleep' not in e.name.lower()]
    t0 = evs[0].time_range.start
    last_end = {}
    print('start_us   dur_us  gap_us stream  kernel')
    for e in evs:
        st = getattr(e, 'stream', None) if hasattr(e, 'stream') else None
        key = st
        gap = e.time_range.start - last_end.get(key, e.time_range.start)
        last_end[key] = e.time_range.end
        print(f'{e.time_range.start - t0:8.1f} {e.time_range.end - e.time_range.start:8.1f} {gap:7.1f} {str(st):>6}  {e.name[:90]}')
    print('frame span us', evs[-1].time_range.end - t0)


if __name__ == '__main__':
    main()

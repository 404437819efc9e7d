import sys, time, torch
sys.path.insert(0, '.')
from srfdet_b200 import synth
from srfdet_b200.pipeline import RegionFeaturePipeline
pipe = RegionFeaturePipeline('nusc', precision='fp16')
pts = torch.as_tensor(synth.cloud('nusc', 1000)).cuda()
for _ in range(5): pipe.run_frame(pts)
torch.cuda.synchronize()
for name, fn in [('frame', lambda: pipe.run_frame(pts)), ('encode', lambda: pipe.encode(pts)), ('stages', lambda: pipe.region_stages())]:
    cpu = []; gpu = []
    for _ in range(10):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); fn(); e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        cpu.append((t1 - t0) * 1e3); gpu.append(e0.elapsed_time(e1))
    cpu.sort(); gpu.sort()
    print(f'{name}: cpu enqueue {cpu[5]:.3f} ms, gpu {gpu[5]:.3f} ms')

"""Per-role cycle breakdown of the tcgen05 sparse-conv kernel, layer by layer.

Needs the instrumented build (python -m srfdet_b200.build --prof) and
SRFDET_B200_LIB=srfdet_b200/csrc/libsrfdet_b200_prof.so.  Development tool: numbers are
clock64() sums per CTA (producer warp 4 lane 0, MMA lane 0, epilogue thread 0)."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srfdet_b200 import _lib as L, synth  # noqa: E402
from srfdet_b200.pipeline import RegionFeaturePipeline  # noqa: E402


def main():
    lib = L.load()
    lib.srf_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
    buf = np.zeros(1024 * 16, dtype=np.uint64)
    tbuf = np.zeros(1024 * 4, dtype=np.uint64)
    lib.srf_prof_read_t.argtypes = [ctypes.c_void_p, ctypes.c_int]

    class Hook(list):
        def append(self, e):
            lib.srf_prof_read(buf.ctypes.data, 1)
            e['prof'] = buf.reshape(1024, 16).copy()
            lib.srf_prof_read_t(tbuf.ctypes.data, lib.srf_prof_read_t(None, -1) - 1)
            e['t'] = tbuf.reshape(1024, 4).copy()
            super().append(e)

    pipe = RegionFeaturePipeline('nusc', fusion=False, precision='fp16')
    enc = pipe.detector.pts_middle_encoder
    enc.overlap_geometry = False
    pts = torch.as_tensor(synth.cloud('nusc', 1)).cuda()
    for _ in range(2):
        pipe.encode(pts)
    torch.cuda.synchronize()
    lib.srf_prof_read(None, 1)
    enc.profile = Hook()
    torch.cuda._sleep(30_000_000)
    pipe.encode(pts)
    torch.cuda.synchronize()
    prof, enc.profile = enc.profile, None
    rows = []
    for e in prof:
        if not e['umma']:
            continue
        p = e['prof']
        g = int(p[0, 10])
        p = p[:g].astype(np.float64)
        ms = e['start'].elapsed_time(e['end'])
        t = e['t'][:g].astype(np.int64)
        t0 = t[:, 0].min()
        f = lambda c: float(np.mean(p[:, c]))
        rows.append(dict(layer=e['layer'], cin=e['cin'], cout=e['cout'], n_out=int(e['n_out']), ms=round(ms, 4), grid=g,
                         kclk_launch=round(ms * 1e-3 * 1.965e9 / 1e3, 1),
                         prod_total=round(f(0) / 1e3, 1), prod_wait_empty=round(f(1) / 1e3, 1), prod_publish=round(f(2) / 1e3, 1),
                         stages_per_cta=round(f(3), 1),
                         mma_total=round(f(4) / 1e3, 1), mma_wait_full=round(f(5) / 1e3, 1), mma_wait_tmem=round(f(6) / 1e3, 1),
                         tiles_per_cta=round(f(7), 2),
                         epi_total=round(f(8) / 1e3, 1), epi_wait_acc=round(f(9) / 1e3, 1),
                         prod_total_max=round(float(p[:, 0].max()) / 1e3, 1),
                         prod_issue=round(f(11) / 1e3, 1), prod_arrive=round(f(12) / 1e3, 1), prod_fetch=round(f(13) / 1e3, 1),
                         mma_issue=round(f(14) / 1e3, 1),
                         span_us=round((t[:, 2].max() - t0) / 1e3, 1), start_last_us=round((t[:, 0].max() - t0) / 1e3, 1),
                         prologue_us=round(float(np.mean(t[:, 1] - t[:, 0])) / 1e3, 1), end_first_us=round((t[:, 2].min() - t0) / 1e3, 1),
                         end_mean_us=round(float(np.mean(t[:, 2] - t0)) / 1e3, 1)))
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/igemm_prof.json', 'w') as fh:
        json.dump(rows, fh, indent=1)
    keys = list(rows[0].keys())
    print(' '.join(f'{k[:10]:>10}' for k in keys))
    for r in rows:
        print(' '.join(f'{str(r[k])[:10]:>10}' for k in keys))


def gaps():
    """Inter-kernel gaps of the tcgen05 launches inside one CUDA-graph replay of the frame."""
    lib = L.load()
    lib.srf_prof_read_t.argtypes = [ctypes.c_void_p, ctypes.c_int]
    tbuf = np.zeros(1024 * 4, dtype=np.uint64)
    pipe = RegionFeaturePipeline('nusc', fusion=False, precision='fp16', use_graph=True)
    pts = torch.as_tensor(synth.cloud('nusc', 1)).cuda()
    for _ in range(3):
        pipe.run_frame(pts)
    torch.cuda.synchronize()
    n = lib.srf_prof_read_t(None, -1)
    # launch ids baked into the graph: the last capture pass issued the last 30 ids (20 convs + 10 linears)
    recs = []
    for lid in range(n - 30, n):
        lib.srf_prof_read_t(tbuf.ctypes.data, lid)
        t = tbuf.reshape(1024, 4).astype(np.int64)
        live = t[:, 0] > 0
        recs.append((lid, int(live.sum()), int(t[live, 0].min()), int(t[live, 2].max())))
    base = recs[0][2]
    prev_end = None
    print('launch  ctas  start_us  span_us  gap_from_prev_end_us')
    for lid, n_cta, t0, t1 in recs:
        gap = (t0 - prev_end) / 1e3 if prev_end is not None else 0.0
        print(f'{lid:6d} {n_cta:5d} {(t0 - base) / 1e3:9.1f} {(t1 - t0) / 1e3:8.1f} {gap:8.1f}')
        prev_end = t1


def trace(dbg_layers=(2, 7, 12, 17)):
    """Event timeline of CTA 0 for a few layers (eager launches)."""
    lib = L.load()
    lib.srf_prof_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.srf_prof_read_t.argtypes = [ctypes.c_void_p, ctypes.c_int]
    ebuf = np.zeros(256, dtype=np.uint64)
    names = {1: 'P publish begin', 2: 'P publish end', 3: 'P slots issued', 4: 'M tile begin', 5: 'M tile committed',
             6: 'E begin', 7: 'E end'}

    class Hook(list):
        def append(self, e):
            lid = lib.srf_prof_read_t(None, -1) - 1
            n = lib.srf_prof_trace(ebuf.ctypes.data, lid)
            e['trace'] = [(int(v) >> 56, int(v) & ((1 << 56) - 1)) for v in ebuf[:n]]
            super().append(e)

    pipe = RegionFeaturePipeline('nusc', fusion=False, precision='fp16')
    enc = pipe.detector.pts_middle_encoder
    enc.overlap_geometry = False
    pts = torch.as_tensor(synth.cloud('nusc', 1)).cuda()
    for _ in range(2):
        pipe.encode(pts)
    torch.cuda.synchronize()
    for lid in range(64):
        lib.srf_prof_trace(ebuf.ctypes.data, lid)
    enc.profile = Hook()
    pipe.encode(pts)
    torch.cuda.synchronize()
    prof, enc.profile = enc.profile, None
    for e in prof:
        if not e['umma'] or e['layer'] not in dbg_layers:
            continue
        tr = sorted(e['trace'], key=lambda x: x[1])
        if not tr:
            continue
        t0 = tr[0][1]
        print(f"--- layer {e['layer']} {e['cin']}>{e['cout']} rows {int(e['n_out'])}")
        for code, t in tr:
            print(f'{(t - t0) / 1e3:9.2f} us  {names.get(code, code)}')


if __name__ == '__main__':
    if 'trace' in sys.argv:
        trace()
        sys.exit(0)
    gaps() if 'gaps' in sys.argv else main()

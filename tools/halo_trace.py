"""Event timeline of CTA 0 of the halo-tile dense conv (development build: python -m srfdet_b200.build --variant trace SRF_HALO_TRACE=1).

    SRFDET_B200_LIB=$PWD/srfdet_b200/csrc/libsrfdet_b200_trace.so python tools/halo_trace.py [h w cin cout]
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from srfdet_b200 import _lib as L   # noqa: E402

NAMES = {1: 'kernel entry', 2: 'prologue done (after griddepcontrol.wait)', 3: 'all roles done', 10: 'epilogue: accumulator ready', 11: 'epilogue: tile stored',
         20: 'halo producer: slot free', 21: 'halo producer: copies issued', 30: 'weight producer: slot free', 40: 'MMA: halo landed',
         41: 'MMA: weight tile landed', 42: 'MMA: tile committed'}

h, w, cin, cout = [int(x) for x in sys.argv[1:5]] if len(sys.argv) >= 5 else (184, 184, 128, 128)
lib = L.load()
raw = ctypes.CDLL(L.SO_PATH)
raw.srf_halo_trace_read.restype = ctypes.c_int
x = torch.randn(h * w, cin, device='cuda').half()
wp = torch.randn(9 * cin * cout, device='cuda').half()
bias = torch.zeros(cout, device='cuda')
y = torch.empty((h * w + 127) // 128 * 128, cout, device='cuda', dtype=torch.float16)
buf = (ctypes.c_ulonglong * 4096)()
for rep in range(3):
    raw.srf_halo_trace_read(buf, 1)
    L.check(lib.srf_conv3x3_rows(L.ptr(x), L.F16, 1, h, w, cin, L.ptr(wp), cout, L.ptr(bias), 1, L.ptr(y), L.F16, L.stream_ptr()), 'conv')
    n = raw.srf_halo_trace_read(buf, 0)
ev = sorted(((buf[i] & 0x00ffffffffffffff, buf[i] >> 56) for i in range(n)))
t0 = ev[0][0]
for t, code in ev:
    print(f'{(t - t0) / 1e3:9.2f} us  {NAMES.get(code, code)}')

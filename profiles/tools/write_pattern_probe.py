"""Write-side ceiling of kernel A: its 13 output streams written with nothing else going on."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from iv_interpolation_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda")
n = 69_093_480
stride = (n + 255) // 256 * 256
buf = torch.empty(13 * stride, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(L.ivx_probe_write_streams(n, buf.data_ptr(), stride, st))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    L.ivx_probe_write_streams(n, buf.data_ptr(), stride, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"13 streams, {n} rows, 93 B/row: {ms:.4f} ms, {n * 93 / ms / 1e6:.1f} GB/s")

"""Dev tool (GPU): tcgen05.mma rate (M128 x N x K16, bf16) when the A window starts off the 1024-byte swizzle atom."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from holocron_b200._lib import lib_path, stream_ptr

L = ctypes.CDLL(str(__import__("pathlib").Path(__file__).resolve().parent / "probes" / "libhb_probes.so"))  # python tools/probes/build.py
L.hb_dev_mma_shift_rate_probe.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p]
torch.zeros(1, device="cuda")
count = 20000
for N in (48, 96, 128):
    for shift, cycle in ((0, 0), (8, 0), (1, 0), (2, 0), (3, 0), (114, 0), (120, 0), (1, 3), (8, 9), (38, 9)):
        L.hb_dev_mma_shift_rate_probe(N, 2000, shift, cycle, 148, stream_ptr()); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.hb_dev_mma_shift_rate_probe(N, count, shift, cycle, 148, stream_ptr()); e1.record(); torch.cuda.synchronize()
        ns = e0.elapsed_time(e1) * 1e6 / count
        print(f"N {N:3d} shift {shift:3d} rows, cycle {cycle}: {ns:6.1f} ns per MMA ({ns * 1.965:5.0f} cycles at 1965 MHz)", flush=True)

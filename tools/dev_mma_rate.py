import sys, ctypes, torch
sys.path.insert(0, ".")
from holocron_b200._lib import lib_path, stream_ptr
L = ctypes.CDLL(str(__import__("pathlib").Path(__file__).resolve().parent / "probes" / "libhb_probes.so"))  # python tools/probes/build.py
L.hb_dev_mma_rate_probe.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p]
torch.zeros(1, device="cuda")
for ctas in (1, 148):
    for N in (16, 32, 48, 64, 96, 128, 192, 256):
        for distinct in (0, 1):
            count = 20000
            L.hb_dev_mma_rate_probe(N, 2000, distinct, ctas, stream_ptr()); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); L.hb_dev_mma_rate_probe(N, count, distinct, ctas, stream_ptr()); e1.record(); torch.cuda.synchronize()
            ns = e0.elapsed_time(e1) * 1e6 / count
            tf = 2 * 128 * N * 16 * ctas / ns / 1e3
            print(f"ctas {ctas:3d} N {N:3d} acc_alt {distinct}: {ns:7.1f} ns per MMA  -> {tf:8.1f} TFLOP/s", flush=True)

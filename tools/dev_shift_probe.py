import sys, ctypes, torch
sys.path.insert(0, ".")
from holocron_b200._lib import lib_path, ptr, stream_ptr
L = ctypes.CDLL(str(__import__("pathlib").Path(__file__).resolve().parent / "probes" / "libhb_probes.so"))  # python tools/probes/build.py
L.hb_dev_umma_shift_probe.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
torch.manual_seed(0)
rows = 384
a = torch.randn(rows, 64, device="cuda").bfloat16()
b = torch.randn(64, 64, device="cuda").bfloat16()
for mode in (0, 1):
    res = []
    for shift in (0, 1, 2, 3, 5, 7, 8, 9, 16, 17, 114, 115, 116, 228, 229, 230, 255):
        out = torch.full((128, 64), float("nan"), device="cuda")
        rc = L.hb_dev_umma_shift_probe(ptr(a), ptr(b), ptr(out), rows, shift, mode, stream_ptr())
        torch.cuda.synchronize()
        ref = a[shift:shift + 128].float() @ b.float().t()
        err = (out - ref).abs().max().item()
        res.append((shift, rc, round(err, 4)))
    print("mode", mode, res, flush=True)

"""Builds the development probes (tcgen05 MMA-rate / shifted-descriptor micro-benchmarks) into their OWN shared object,
tools/probes/libhb_probes.so - they are not part of the product library holocron_b200/csrc/libholocron_b200.so."""
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE.parent.parent / "holocron_b200" / "csrc"
OUT = HERE / "libhb_probes.so"


def build() -> Path:
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", str(CSRC), "-shared", "-cudart", "static",
           "-o", str(OUT), str(HERE / "dev_probe.cu"), str(CSRC / "runtime.cu"), "-ldl"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build())
    sys.exit(0)

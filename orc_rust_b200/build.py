"""Builds orc_rust_b200/liborc_b200.so in-tree with nvcc for sm_100a (B200).

    python -m orc_rust_b200.build [--force]

The shared library has a plain C ABI (include/orc_b200.h); cudart is linked statically so the
library loads on a machine without a GPU (host-only calls work; CUDA calls report ORCB_CUDA)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liborc_b200.so")
SOURCES = ["k_int.cu", "k_streams.cu", "k_strings.cu", "k_decompress.cu", "meta.cc", "tz.cc", "schema.cc", "plan.cc", "job.cc", "export.cc", "selection.cc", "predicate.cc", "c_api.cc"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-x", "cu",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "orc_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    extra = os.environ.get("ORCB_NVCC_DEFS", "").split()  # e.g. "-DORCB_LZ_SB=32 -DORCB_LZ_CTAS=7" (experiments)
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src + ".o")
        objs.append(obj)
        cmd = ["nvcc"] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
        if verbose:
            sys.stderr.write(out.decode())
    cmd = ["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

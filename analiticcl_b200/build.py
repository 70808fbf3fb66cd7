"""Builds libanaliticcl_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m analiticcl_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libanaliticcl_b200.so")
SOURCES = ["kernels.cu", "export.cu", "engine.cu", "shard_comm.cu", "gpu_build.cu", "gpu_segment.cu", "host_model.cpp", "editscript.cpp", "search.cpp", "sequence.cpp", "capi.cpp"]
HEADERS = ["shard_comm.cu", "device_types.h", "editscript_fixed.h", "kernel_common.cuh", "kernels.h", "engine.h", "host_model.h", "hostpool.h", "search.h", "unicode_tables.h",
           os.path.join("..", "..", "include", "analiticcl_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-ccbin", "/usr/bin/g++",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function,-Wno-attributes,-pthread",
         "--cudart", "static"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    extra = os.environ.get("ANL_NVCC_EXTRA", "").split()  # developer knob: extra nvcc flags (e.g. -DANL_K1_WARPS=4)
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, "build", s + ".o")
        objs.append(o)
        cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose and s.endswith(".cu") else []) + \
              ["-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {s}\n{out}\n")
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC, "-shared", "-Wno-deprecated-gpu-targets", "-o", OUT] + objs + ["--cudart", "static", "-ccbin", "/usr/bin/g++", "-Xcompiler", "-pthread"]
    subprocess.check_call(cmd + ["-ldl"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

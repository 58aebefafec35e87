"""Build libfdn_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI).

    python -m fdn_tip2025_b200.build            # product library  -> fdn_tip2025_b200/libfdn_b200.so
    python -m fdn_tip2025_b200.build --emu      # host emulation   -> tests/emu/_build/libfdn_emu.so (debug only)
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "fft_global.cu", "patch_spectral.cu", "pointwise.cu", "pw_mma.cu", "conv.cu", "conv_mma.cu", "lpnet.cu", "imgio.cu", "metrics.cu", "losses.cu"]
LIB = os.path.join(HERE, "libfdn_b200.so")
EMU_DIR = os.path.join(ROOT, "tests", "emu", "_build")
EMU_LIB = os.path.join(EMU_DIR, "libfdn_emu.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "--use_fast_math=false"]


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    return h.hexdigest()


def _all_inputs():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    return files


def build(verbose=False, force=False):
    """Compile every CUDA source for sm_100a into one shared library.  Returns the library path."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    stamp = LIB + ".sha256"
    digest = _digest(_all_inputs(), " ".join(NVCC_FLAGS) + ",".join(SOURCES) + os.environ.get("FDN_MMA_PROFILE", ""))
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    if os.environ.get("FDN_MMA_PROFILE") == "1":       # dev build: per-role wait counters in k_pw_mma (tools/mma_wait_profile.py)
        flags.append("-DFDN_MMA_PROFILE=1")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print("---- %s\n%s" % (src, out), flush=True)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-shared", "-o", LIB, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


def build_emu(force=False):
    """g++ build of the same kernel sources on the CUDA-semantics emulator (tests/emu).  Debug aid only."""
    os.makedirs(EMU_DIR, exist_ok=True)
    emu_src = os.path.join(ROOT, "tests", "emu", "cuda_emu.cpp")
    inputs = _all_inputs() + [emu_src, os.path.join(ROOT, "tests", "emu", "cuda_emu.h")]
    stamp = EMU_LIB + ".sha256"
    digest = _digest(inputs)
    if not force and os.path.exists(EMU_LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return EMU_LIB
    base = ["g++", "-std=c++20", "-O2", "-fPIC", "-fvisibility=hidden", "-DFDN_EMU", "-I", os.path.join(ROOT, "tests", "emu"),
            "-I", CSRC, "-fno-fast-math", "-ffp-contract=off", "-Wno-unused-result"]
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(EMU_DIR, src.replace(".cu", ".o"))
        procs.append(subprocess.Popen(base + ["-x", "c++", "-c", os.path.join(CSRC, src), "-o", obj]))
        objs.append(obj)
    obj = os.path.join(EMU_DIR, "cuda_emu.o")
    procs.append(subprocess.Popen(base + ["-c", emu_src, "-o", obj]))
    objs.append(obj)
    if any(p.wait() != 0 for p in procs):
        raise RuntimeError("g++ (emulation build) failed")
    subprocess.check_call(["g++", "-shared", "-o", EMU_LIB, *objs, "-lpthread"])
    with open(stamp, "w") as f:
        f.write(digest)
    return EMU_LIB


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu(force="--force" in sys.argv))
    else:
        print(build(verbose="-v" in sys.argv, force="--force" in sys.argv))

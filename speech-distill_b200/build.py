"""Build recipe for libkd_b200.so (hand-written CUDA for sm_100a behind a C ABI).

    python speech-distill_b200/build.py [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with gpurun snapshots.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkd_b200.so")
SOURCES = ["kd_api.cu", "kd_stream.cu", "kd_topk.cu", "kd_rows.cu", "kd_probe.cu", "kd_multimem.cu", "kd_fused.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--split-compile", "0",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC,-O3",
    "-Xptxas", "-v",
    "-DKD_BUILDING_LIB",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "kd_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, defines=()):
    """variant=NAME builds libkd_b200_NAME.so from the same sources with extra -D defines (same-box A/B runs through
    KD_B200_LIB); objects go to build/NAME/."""
    lib = LIB if variant is None else os.path.join(HERE, f"libkd_b200_{variant}.so")
    if variant is None and not force and not _stale():
        return LIB
    objs = []
    bdir = os.path.join(HERE, "build") if variant is None else os.path.join(HERE, "build", variant)
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, *defines, "-c", path, "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    link = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append("==== link\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("link failed")
    with open(os.path.join(bdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return lib


if __name__ == "__main__":
    variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=variant,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")]))

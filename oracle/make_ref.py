"""Recipe for oracle/_ref/: the UNMODIFIED reference loss module, copied from where it lies under /root/reference.

    python oracle/make_ref.py

The reference is pure Python (`distillation_loss.py` imports only torch), so "building" it is a byte-for-byte
copy.  Outputs go to oracle/_ref/ only (git-ignored, NOT gpurun-ignored: the copy travels to the GPU box, where
/root/reference does not exist).  It is test / measurement infrastructure: tests/ pins the oracle restatement
against it and bench.py times it on the host cores as the CPU baseline (cpu_baseline.kind = "reference") and, on
the GPU, as the eager baseline (tools/gpu_baselines.py).  Nothing under speech-distill_b200/ imports it.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("KD_REFERENCE_ROOT", "/root/reference")
FILES = ["distillation_loss.py"]  # reference distillation_loss.py:1-128, the whole arithmetic of the path


def make():
    out = os.path.join(HERE, "_ref")
    if not os.path.isdir(REF):
        return None  # GPU box: use what the snapshot brought along
    os.makedirs(out, exist_ok=True)
    lines = []
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(out, f)
        shutil.copyfile(src, dst)
        lines.append(f"{hashlib.sha256(open(dst, 'rb').read()).hexdigest()}  {f}")
    with open(os.path.join(out, "SOURCE.txt"), "w") as fh:
        fh.write("byte-for-byte copies made by oracle/make_ref.py from the reference checkout\n" + "\n".join(lines) + "\n")
    return out


def load_reference_module():
    """The unmodified reference module (oracle/_ref/distillation_loss.py) or None if the recipe has not run."""
    import importlib.util

    path = os.path.join(HERE, "_ref", "distillation_loss.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("reference_distillation_loss", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(make() or f"{REF} not present; nothing done", file=sys.stderr)

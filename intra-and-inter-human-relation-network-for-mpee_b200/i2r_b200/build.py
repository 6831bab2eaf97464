"""Build csrc/*.cu into i2r_b200/libi2r_sm100.so with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
LIB_PATH = os.path.join(HERE, "libi2r_sm100.so")
SOURCES = ["igemm_tc.cu", "conv_halo.cu", "misc_kernels.cu", "attention.cu", "attention_tc.cu", "window_attention_tc.cu", "encoder_tail.cu", "stem_tc.cu", "hrformer_kernels.cu", "postproc.cu", "preproc.cu", "mask_res.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newest_mtime(paths):
    return max(os.path.getmtime(p) for p in paths)


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    srcs = [os.path.join(CSRC, s) for s in os.listdir(CSRC)]
    srcs.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "i2r.h"))
    return _newest_mtime(srcs) > os.path.getmtime(LIB_PATH)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + [
        os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

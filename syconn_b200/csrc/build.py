"""Build libsyk.so (sm_100a) in-tree with nvcc.  `python -m syconn_b200.csrc.build [--force]`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, os.environ.get("SYK_LIB_NAME", "libsyk.so"))
SOURCES = ["syk_table.cu", "syk_props.cu", "syk_cs.cu", "syk_morph.cu", "syk_host.cu", "syk_lz4.cu", "syk_ccl.cu", "syk_morph_vol.cu"]
HEADERS = ["syk_common.cuh", "syk_cs_fast.cuh", "syk_cs_march.cuh", os.path.join("..", "..", "include", "syk.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
         "--use_fast_math", "-Xptxas", "-v"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + os.environ.get("SYK_NVCC_EXTRA", "").split() + ["-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed for " + src)
        with open(obj + ".ptxas.txt", "w") as f:
            f.write(r.stderr)
        objs.append(obj)
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-lcudart_static", "-lrt", "-lpthread", "-ldl"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Builds libcpml_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels
to the GPU box with the repo snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["cpml_api.cu", "kernels_3d.cu", "kernels_3d_tma.cu", "kernels_3d_visco.cu", "kernels_2d.cu", "kernels_2d_visco.cu", "cpml_host.cpp", "attenuation_fit.cpp"]
HEADERS = [os.path.join(CSRC, "cpml_internal.h"),
           os.path.join(HERE, "..", "include", "cpml_b200.h")]
LIB = os.path.join(HERE, "libcpml_b200.so")

# -fmad=false: products and sums are rounded separately, in source order, so the fields
# are bit-identical to an IEEE (non-FMA) build of the reference loops.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", "-cudart", "static"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libcpml_b200.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None,
          out: str | None = None) -> str:
    target = out or LIB
    if not force and out is None and not needs_build():
        return target
    cmd = [nvcc_path()] + NVCC_FLAGS + (extra_flags or [])
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", target] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports CC=/opt/gcc/bin/gcc; let nvcc pick its own host compiler
    env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""Builds libcpml_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels
to the GPU box with the repo snapshot).  One object per source, compiled in parallel and
rebuilt only when the source (or a header) is newer; `force=True` rebuilds everything."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(CSRC, "_build")
SOURCES = ["cpml_api.cu", "cpml_multi.cu", "kernels_3d.cu", "kernels_3d_tma.cu", "kernels_3d_ws.cu", "kernels_3d_visco.cu", "kernels_3d_visco_ws.cu",
           "kernels_2d.cu", "kernels_2d_ws.cu", "kernels_2d_visco.cu", "cpml_host.cpp", "attenuation_fit.cpp"]
HEADERS = [os.path.join(CSRC, "cpml_internal.h"), os.path.join(CSRC, "tma_common.cuh"), os.path.join(CSRC, "visco_common.cuh"), os.path.join(CSRC, "kernels_2d_point.cuh"),
           os.path.join(HERE, "..", "include", "cpml_b200.h")]
LIB = os.path.join(HERE, "libcpml_b200.so")

# -fmad=false: products and sums are rounded separately, in source order, so the fields
# are bit-identical to an IEEE (non-FMA) build of the reference loops.  The host
# side gets -ffp-contract=off for the same reason (gcc contracts to FMA by default on aarch64).
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-fmad=false",
                     "-Xcompiler", "-fPIC,-O2,-Wall,-ffp-contract=off"]
LINK_FLAGS = ARCH + ["-shared", "-cudart", "static"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libcpml_b200.so cannot be built")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _obj(src: str, objdir: str) -> str:
    return os.path.join(objdir, os.path.splitext(src)[0] + ".o")


def _stale(src: str, objdir: str) -> bool:
    o = _obj(src, objdir)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    deps = [os.path.join(CSRC, src)] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in _sources()] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _env():
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports CC=/opt/gcc/bin/gcc; let nvcc pick its own host compiler
    env.pop("CXX", None)
    return env


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True, env=_env())
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stderr


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None,
          out: str | None = None) -> str:
    """extra_flags / out: a variant build (A/B measurements) into its own object directory."""
    target = out or LIB
    if not force and out is None and not extra_flags and not needs_build():
        return target
    objdir = OBJDIR if out is None and not extra_flags else OBJDIR + "_" + os.path.basename(target)
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    flags = NVCC_FLAGS + (extra_flags or []) + (["-Xptxas", "-v"] if verbose else [])
    todo = [s for s in _sources() if force or _stale(s, objdir)]
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        list(ex.map(lambda s: _run([nvcc] + flags + ["-c", "-o", _obj(s, objdir), os.path.join(CSRC, s)], verbose), todo))
    _run([nvcc] + LINK_FLAGS + ["-o", target] + [_obj(s, objdir) for s in _sources()], verbose)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

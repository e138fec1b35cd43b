"""Build the C-ABI shared library (libpli_attention.so) for sm_100a with nvcc.

`python -m physics_llm_inference_b200.build` (or `__graft_entry__.build()`) compiles every
`csrc/*.cu` with `-gencode arch=compute_100a,code=sm_100a -lineinfo` and links them into
`physics_llm_inference_b200/libpli_attention.so`, in-tree, so the library travels with the repo
snapshot.  nvcc cross-compiles without a GPU.  Objects are rebuilt only when a source or header is
newer than the object.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libpli_attention.so")

SOURCES = ["pli_capi.cu", "prefill_tcgen05.cu", "prefill_simt.cu", "decode.cu", "online_softmax.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(ROOT, "include", "pli_attention.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libpli_attention.so")
    return nvcc


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(nvcc: str, src: str, obj: str, verbose: bool, defines=()) -> str:
    cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    log = r.stdout + r.stderr
    with open(obj + ".ptxas.log", "w") as f:
        f.write(log)
    if verbose:
        sys.stderr.write(log)
    return log


def build(force: bool = False, verbose: bool = False, variant: str = "", defines=()) -> str:
    """Build the library.  `variant` + `defines` produce a tuning build next to the product one
    (build/libpli_attention_<variant>.so, e.g. -DPLI_PROFILE=1 or -DPLI_POLY_PAIRS=6); the product
    library is always the default, define-free build."""
    nvcc = find_nvcc()
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    objs = []
    suffix = f"_{variant}" if variant else ""
    lib = os.path.join(OBJ, f"libpli_attention{suffix}.so") if variant else LIB
    for name in SOURCES:
        src = os.path.join(CSRC, name)
        obj = os.path.join(OBJ, name.replace(".cu", f"{suffix}.o"))
        objs.append(obj)
        if force or _stale(obj, [src, *HEADERS]):
            jobs.append((src, obj))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(5, len(jobs))) as ex:
            list(ex.map(lambda so: _compile(nvcc, so[0], so[1], verbose, tuple(defines)), jobs))
    if force or jobs or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    variant = ""
    defines = [a for a in sys.argv[1:] if a.startswith("-D")]
    for a in sys.argv[1:]:
        if a.startswith("--variant="):
            variant = a.split("=", 1)[1]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=variant, defines=defines))

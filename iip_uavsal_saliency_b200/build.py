"""In-tree build of libuavsal_b200.so (nvcc, sm_100a only).  Cross-compiles without a GPU.

Every csrc/*.cu is compiled to its own object (in parallel, cached on a digest of the source, the headers and the flags)
under csrc/.obj/, then linked; an unchanged tree is a no-op."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, ".obj")
LIB = os.path.join(PKG, "libuavsal_b200.so")
STAMP = LIB + ".stamp"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _headers_digest() -> bytes:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    files.append(os.path.join(os.path.dirname(PKG), "include", "uavsal_b200.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.digest()


def _src_digest(src: str, hdr: bytes) -> str:
    with open(os.path.join(CSRC, src), "rb") as fh:
        return hashlib.sha256(hdr + fh.read()).hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libuavsal_b200.so next to this file; no-op when sources are unchanged."""
    hdr = _headers_digest()
    srcs = sources()
    digs = {s: _src_digest(s, hdr) for s in srcs}
    total = hashlib.sha256("".join(s + digs[s] for s in srcs).encode()).hexdigest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == total:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: str):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        stamp = obj + ".stamp"
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == digs[src]:
            return None
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            return "nvcc failed on %s:\n%s%s" % (src, proc.stdout, proc.stderr)
        with open(stamp, "w") as fh:
            fh.write(digs[src])
        return None

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        errors = [e for e in ex.map(compile_one, srcs) if e]
    if errors:
        raise RuntimeError("\n".join(errors))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    with open(STAMP, "w") as fh:
        fh.write(total)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

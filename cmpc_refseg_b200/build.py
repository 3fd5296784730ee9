"""Builds libcmpc_b200.so (sm_100a only) in-tree with nvcc.  No GPU needed: nvcc cross-compiles."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libcmpc_b200.so"
STAMP = PKG / ".libcmpc_b200.stamp"
OBJ = PKG / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math" if False else "-DCMPC_NO_FAST_MATH",   # accuracy matters more than MUFU shortcuts
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "cmpc_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _fresh(dig: str) -> bool:
    return LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == dig


def build(force: bool = False, verbose: bool = False) -> Path:
    dig = _digest()
    if not force and _fresh(dig):
        return LIB
    # several ranks of one torchrun may get here at once: one of them compiles (to a temporary file, renamed into place), the
    # others wait on the lock and find the library fresh
    import fcntl
    with open(PKG / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _fresh(dig):
                return LIB
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            tmp = LIB.with_suffix(f".tmp{os.getpid()}.so")
            # one object per source, compiled in parallel and cached by content (headers + flags included): an edit to one
            # kernel recompiles one file
            from concurrent.futures import ThreadPoolExecutor
            OBJ.mkdir(exist_ok=True)
            hdr = hashlib.sha256()
            for f in sorted(list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "cmpc_b200.h"]):
                hdr.update(f.read_bytes())
            hdr.update(" ".join(NVCC_FLAGS).encode())
            cflags = [x for x in NVCC_FLAGS if x != "-shared"]

            def compile_one(src):
                h = hashlib.sha256(hdr.digest() + src.read_bytes()).hexdigest()[:16]
                obj = OBJ / f"{src.stem}.{h}.o"
                if obj.exists() and not force:
                    return obj, ""
                for old in OBJ.glob(f"{src.stem}.*.o"):
                    old.unlink(missing_ok=True)
                cmd = [nvcc, *cflags, *( ["-Xptxas", "-v"] if verbose else []), "-c", "-o", str(obj), str(src)]
                r = subprocess.run(cmd, capture_output=True, text=True)
                if r.returncode != 0:
                    obj.unlink(missing_ok=True)
                    raise RuntimeError(f"nvcc failed on {src.name}:\n" + r.stdout + r.stderr)
                return obj, r.stderr
            with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
                res = list(ex.map(compile_one, _sources()))
            if verbose:
                for _, err in res:
                    print(err, file=sys.stderr)
            r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", str(tmp),
                                *[str(o) for o, _ in res]], capture_output=True, text=True)
            if r.returncode != 0:
                tmp.unlink(missing_ok=True)
                raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
            os.replace(tmp, LIB)
            STAMP.write_text(dig)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

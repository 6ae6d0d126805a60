"""Build libcurdle_b200.so in-tree with nvcc for sm_100a (no JIT cache: the
built .so travels to the GPU box with the repo snapshot)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcurdle_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # host side (Merlin / Keccak, Fr arithmetic between launches): BMI1/2 give rorx / andn / mulx,
    # worth ~25 % on the transcript and Fr code; every x86-64 host since 2013 (Haswell, Zen) has them
    "-Xcompiler", "-fPIC,-O3,-mbmi,-mbmi2", "-Xptxas", "-v",
]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp")):
                out.append(os.path.join(root, f))
    out.append(os.path.join(HERE, "..", "include", "curdle_b200.h"))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def _compile_one(args):
    src, obj = args
    cmd = ["nvcc", "-c"] + NVCC_FLAGS + ["-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res.returncode, " ".join(cmd) + "\n" + res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every translation unit (in parallel) and link libcurdle_b200.so."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    units = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    host_dir = os.path.join(CSRC, "host")
    if os.path.isdir(host_dir):
        units += [os.path.join(host_dir, f) for f in sorted(os.listdir(host_dir)) if f.endswith((".cpp", ".cu"))]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdr_mtime = max(os.path.getmtime(s) for s in _sources() if not s.endswith((".cu", ".cpp")))
    jobs = []
    objs = []
    for u in units:
        obj = os.path.join(objdir, os.path.basename(u) + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(u), hdr_mtime):
            jobs.append((u, obj))
    log = os.path.join(HERE, "build.log")
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex, open(log, "w") as fh:
        for src, rc, out in ex.map(_compile_one, jobs):
            fh.write(out)
            if verbose or rc != 0:
                sys.stderr.write(out)
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {src} (see {log})")
    cmd = ["nvcc", "-shared", "-o", LIB] + objs + ["-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""The product's host-side Fiat-Shamir code (csrc/host/: Keccak / STROBE / Merlin, common.Rand, Fr arithmetic,
the thread pool, the eight-way fiber hashing) under the compiler's sanitizers.  tests/hostcheck/host_sanitize.cpp
builds those sources without nvcc and repeats the library's own self-tests (Merlin known answer, Fr identities,
per-proof transcripts on the pool as fibers == the same transcripts computed one by one); here it runs under
AddressSanitizer + UndefinedBehaviorSanitizer and under ThreadSanitizer, with and without the fibers."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
HOST = os.path.join(HERE, "..", "go-curdleproofs_b200", "csrc", "host")


def build(tmp, name, san):
    out = str(tmp / name)
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-mbmi", "-mbmi2", f"-fsanitize={san}", "-fno-sanitize-recover=all", "-o", out,
           os.path.join(HERE, "hostcheck", "host_sanitize.cpp"), os.path.join(HOST, "fiber.cpp"),
           os.path.join(HOST, "keccak_x8.cpp"), "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0 and ("cannot find" in res.stderr or "unrecognized" in res.stderr):
        pytest.skip(f"g++ has no -fsanitize={san} runtime here")
    assert res.returncode == 0, res.stderr
    return out


@pytest.mark.parametrize("san", ["address,undefined", "thread"])
def test_host_code_is_clean_under_sanitizers(tmp_path, san):
    exe = build(tmp_path, "host_" + san.split(",")[0], san)
    for no_fibers in ("0", "1"):
        env = dict(os.environ, ASAN_OPTIONS="detect_stack_use_after_return=0", TSAN_OPTIONS="halt_on_error=1")
        env.pop("CDL_NO_FIBERS", None)
        if no_fibers == "1":
            env["CDL_NO_FIBERS"] = "1"
        res = subprocess.run([exe, "37", "12"], capture_output=True, text=True, env=env, timeout=300)
        assert res.returncode == 0 and res.stdout.startswith("ok"), res.stdout + res.stderr
        assert "runtime error" not in res.stderr and "Sanitizer" not in res.stderr, res.stderr


def test_field_and_curve_headers_are_clean_under_ubsan():
    """tests/test_hostcheck.py (the product's csrc/*.cuh field / group / GLV / batch-affine / fixed-base code compiled
    for the host with bit-exact carry emulation) once more with UndefinedBehaviorSanitizer: shifts, signed overflow,
    misaligned or out-of-bounds accesses abort the run."""
    import sys

    env = dict(os.environ, CDL_HOSTCHECK_FLAGS="-fsanitize=undefined -fno-sanitize-recover=all -static-libubsan")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_hostcheck.py"), "-x", "-q",
                          "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, timeout=900)
    if "cannot find" in res.stdout + res.stderr:
        pytest.skip("no static libubsan here")
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]

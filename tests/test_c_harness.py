"""include/udales_gpu.h from a C translation unit (gcc -std=c11 -pedantic -Werror): struct layout == the ctypes mirror,
the library links and refuses to run without a GPU; with a GPU the C program runs init / substep / finalize itself."""
import ctypes as C
import json
import os
import subprocess

import pytest

import udales_b200 as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    exe = os.path.join(tmp_path, "harness")
    lib_dir = os.path.dirname(U.LIB_PATH)
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cmd = [cc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi", "harness.c"), "-o", exe, "-L", lib_dir, "-ludales_gpu", f"-Wl,-rpath,{lib_dir}", "-lm"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_struct_layout_matches_ctypes(tmp_path):
    exe = build(str(tmp_path))
    out = subprocess.run([exe, "layout"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lay = json.loads(out.stdout)
    assert lay["sizeof"] == C.sizeof(U.Cfg)
    assert lay["abi"] == U.ABI_VERSION
    assert lay["nfields"] == len(U.FIELD_IDS)
    for name, off in lay.items():
        if name in ("sizeof", "abi", "nfields"):
            continue
        assert getattr(U.Cfg, name).offset == off, name


def test_c_program_without_gpu_gets_enodev(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = build(str(tmp_path))
    out = subprocess.run([exe, "run"], capture_output=True, text=True)
    assert out.returncode == 3 and "no CPU fallback" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_c_program_drives_the_library(tmp_path):
    exe = build(str(tmp_path))
    out = subprocess.run([exe, "run"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok rk3step 3")

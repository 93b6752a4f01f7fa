"""examples/host_lsqr.c: a plain-C host of the C ABI (what a compiled-language driver links against). Without a GPU it
must fail loudly at tfx_init (no CPU fallback); on a GPU box it solves the diagonal system of the reference's solver
unit test (src/tests/tests_lsqr.f90:71-118)."""
import os
import subprocess

import pytest

import tomofastx_b200 as tfx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    tfx.build()
    exe = str(tmp_path / "host_lsqr")
    libdir = os.path.dirname(tfx.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "host_lsqr.c"), "-o", exe, "-L", libdir, "-ltfx",
                           "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"])
    return exe


def _has_gpu():
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60).stdout.count("GPU ") > 0
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="CPU-only check")
def test_c_host_fails_loudly_without_gpu(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device available" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_host_solves_on_gpu(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    words = r.stdout.split()
    x = [float(w) for w in words[2:5]]
    assert all(abs(v - 1.0) < 1e-10 for v in x), r.stdout

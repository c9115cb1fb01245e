"""include/abo.h from a non-Python host: tests/abi_smoke.c (C99, links libabo_cuda.so only) is compiled with gcc against
the public header and run.  Without a GPU the program must stop at abo_ctx_create with exit code 77 (no CPU fallback);
on a B200 it must run the whole create -> fit -> sweep -> append -> clone -> NLML -> GradientGP -> destroy sequence and
reproduce the reference's closed-form known answers."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "abstractbayesopt.jl_b200")


def _build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "abi_smoke")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi_smoke.c"),
           "-o", exe, "-L", PKG, "-labo_cuda", f"-Wl,-rpath,{PKG}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_abi_smoke_compiles_links_and_refuses_cpu(tmp_path):
    exe = _build(tmp_path)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")          # no device visible: the library must refuse, not fall back
    r = subprocess.run([exe], capture_output=True, text=True, env=env)
    assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_abi_smoke_runs_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "abi_smoke ok" in r.stdout

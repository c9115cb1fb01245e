"""Multi-rank checks of the NCCL paths on real GPUs: skipped on a single-GPU box, the gloo world-2 test in
test_parallel_cpu.py covers the host logic there.  Each test launches one process per GPU through torch.distributed.run
(rendezvous on 127.0.0.1) and asserts on the tool's exit code; the tools themselves compare against the oracle / the
single-GPU result (tools/nccl_check.py) and bit for bit between iterations (tools/multi_gpu_stress.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc, script, *args, timeout=420):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, script), *args]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_nccl_paths_two_ranks():
    """abo_gp_sync gives bit-identical posteriors on every rank; sharded top-k and sharded NLML equal the single-GPU
    results; collective failure outcomes leave the communicator usable."""
    r = _torchrun(2, "tools/nccl_check.py")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "nccl_check ok" in r.stdout


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_bo_iteration_is_deterministic_across_ranks_and_iterations():
    """The sharded BO iteration of bench.py (append, posterior broadcast, sweep, top-k all-gather) repeated on identical
    inputs: every score, the local and the global top-100 identical bit for bit in every iteration on every rank."""
    n = min(_ngpu(), 4)
    r = _torchrun(n, "tools/multi_gpu_stress.py", "--nobs", "2048", "--cands", "131072", "--iters", "12")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    rep = json.loads(line)
    assert rep["ranks"] == n and rep["deterministic"] is True, rep

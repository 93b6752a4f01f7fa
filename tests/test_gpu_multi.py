"""Two ranks, one GPU each, NCCL inside libtfx: column-split lsqr_solve_sensit (fused dense sweep and the
compressed split path) against the single-rank oracle. Skipped on a one-GPU box (run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import subprocess

import pytest

from tests.test_dist_model import run_case

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_two_rank_nccl_solve_matches_oracle():
    r = run_case("nccl", 2)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("multi_rank_case ok") == 5, r.stdout
    assert "kind=dist_sensit" in r.stdout

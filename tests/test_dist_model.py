"""world_size-2 run of the column-split LSQR decomposition on CPU (torch.distributed gloo): the partition
helpers, the slab builders and the reduction structure of csrc/lsqr.cu's multi-rank path, checked against
the single-rank oracle solve. The GPU/NCCL version of the same case is tests/test_gpu_multi.py."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def run_case(backend, world, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "multi_rank_case.py"), "--backend", backend]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("world", [2, 3])
def test_column_split_model_matches_oracle(world):
    r = run_case("model", world)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    # five solves on rank 0; the re-partitioning model and the distributed-wavelet model on every rank
    assert r.stdout.count("multi_rank_case ok") == 5 + 2 * world, r.stdout

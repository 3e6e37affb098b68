"""Multi-rank correctness under the driver: when the box shows >= 2 GPUs, tests/multi_gpu_check.py runs under torch.distributed.run
(one process per GPU): every rank's slice of the sharded ensemble must equal the unsharded run BITWISE (paths, accept decisions),
fetch_ll / accept counts through dmt_allreduce_stats (NCCL communicator and the library's own peer-memory kernel) must equal the
unsharded sums, and the all-reduced values must be identical on every rank.  On a 1-GPU box the test is skipped (the sharding
arithmetic itself is covered on CPU by tests/test_multiprocess_gloo.py and, single-GPU, by test_gpu_parity's sharding test)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.own_lanes]


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_ensemble_equals_unsharded_run(world):
    n = _n_gpus()
    if n < world:
        pytest.skip("needs %d GPUs, this box shows %d" % (world, n))
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "PASS" in r.stdout, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])

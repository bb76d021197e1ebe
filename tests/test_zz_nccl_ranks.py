"""The real multi-rank path (-m gpu, needs at least two GPUs): one process per GPU under torchrun, the library's NCCL binding
behind the all-reduce hook, compared with a single context -- see tests/nccl_worker.py.  Skipped on a one-GPU box (NCCL does
not accept two ranks on one device); run with `gpurun --gpus 2` (log in profiles/)."""
import os
import subprocess
import sys

import pytest

from pingpongpro_b200 import api

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ranks", [2, 4, 8])
def test_nccl_ranks_equal_single_context(ranks):
    if api.device_count() < ranks:
        pytest.skip(f"needs {ranks} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1", "--master-port", str(29500 + ranks),
           os.path.join(ROOT, "tests", "nccl_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(p.stdout[-4000:])
    assert p.returncode == 0, p.stderr[-4000:]
    assert p.stdout.count("identical") >= 4

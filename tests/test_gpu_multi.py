"""N > 1: frame-sharded solve with the NCCL allreduce of the reduced system equals the single-problem CPU oracle
(same accept/reject sequence, final RMS within 1e-6 px, parameters within 1e-6). Needs >= 2 GPUs on the box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [2])
def test_sharded_solve_matches_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(HERE, "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("MGPU_OK") == 3 + 7   # 3 default-option solves + 7 LM branch cases

"""Multi-GPU parity (needs >= 2 GPUs, e.g. `gpurun --gpus 2`): the row-sharded
PCG reproduces the single-GPU solution and the optimiser loop stays identical
on every rank."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_solve_and_loop(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + world),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(res.stdout[-3000:], res.stderr[-3000:])
    assert res.returncode == 0
    assert "DIST_OK" in res.stdout

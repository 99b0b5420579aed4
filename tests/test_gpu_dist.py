"""Multi-rank parity of the sharded paths: the z-slab-sharded PCG (matrix-free
level 0, sharded multigrid levels, sharded Helmholtz filter) reproduces the
single-GPU solution and the optimiser loop stays identical on every rank.

With enough GPUs the ranks run one per GPU over NCCL; on a one-GPU box the same
N-rank job runs on cuda:0 with the host shared-memory transport
(csrc/comm.cuh), so the sharded code is exercised wherever the suite runs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,shard_min", [(2, 300), (3, 300), (4, 1500), (2, 10**9)])
def test_sharded_solve_and_loop(world, shard_min):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ)
    env["SKTOPT_B200_MG_SHARD_MIN"] = str(shard_min)
    if torch.cuda.device_count() < world:
        env["SKTOPT_DIST_ONE_GPU"] = "1"
    elif shard_min == 300:
        env["SKTOPT_B200_P2P_HALO"] = "1"     # one GPU per rank: cover the peer-memory halos too
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + world + (1 if shard_min > 10**6 else 0) * 10),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    print(res.stdout[-3000:], res.stderr[-3000:])
    assert res.returncode == 0
    assert "DIST_OK" in res.stdout


def test_sharded_level_with_single_precision_values():
    """A z-slab-sharded multigrid level that is large enough for the bulk-async SpMV
    (>= 8000 rows per rank) streams its fp32 copy in the V-cycle, like C5's level 1 on
    several GPUs: 73 x 55 x 37 nodes, level 1 = 19,684 nodes on 2 ranks."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ)
    env.update(SKTOPT_B200_MG_SHARD_MIN="300", SKTOPT_DIST_MESH="0.11",
               SKTOPT_B200_MG_FP32_MIN_NODES="5000")
    if torch.cuda.device_count() < 2:
        env["SKTOPT_DIST_ONE_GPU"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    print(res.stdout[-3000:], res.stderr[-3000:])
    assert res.returncode == 0
    assert "DIST_OK" in res.stdout
    assert "fp32_sharded_levels=1" in res.stdout

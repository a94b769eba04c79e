"""ONE proof sharded over several ranks (csrc/comm.cu): commitment, Fiat-Shamir challenges and proof bytes must
equal the CPU oracle's — i.e. the single-GPU bytes.

The sharded path's exchanges (tagged stores into every process's shared pinned host segment per round, all-gathers as
peer-memory stores through CUDA IPC) do not need one DEVICE per rank, so the check runs on a single-GPU box too: 2 and
4 ranks time-slicing GPU 0 (gloo carries the job id).  With >= 2 GPUs visible the same check also runs one rank per GPU
(NCCL for the plumbing, P2P over NVLink for the exchanges)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, same_gpu, extra_env=None, timeout=1500):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ)
    if same_gpu:
        env["LASSO_SHARD_SAME_GPU"] = "1"
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "sharded_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert "SHARDED_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_sharded_two_ranks_one_gpu_bit_exact():
    _run(2, True)


def test_sharded_four_ranks_one_gpu_bit_exact():
    _run(4, True)


def test_sharded_two_ranks_one_gpu_no_tables_bit_exact():
    # the bucket-MSM / per-step opening path of a sharded proof (no digit-multiples tables)
    _run(2, True, {"LASSO_B200_NO_MULTIPLES": "1"})


def test_sharded_two_gpus_bit_exact():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (the one-GPU variants above cover the same code on this box)")
    _run(2, False)

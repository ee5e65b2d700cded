"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on the 1-GPU box): the processes are launched with torchrun."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_xjoin_over_cuda_ipc_two_processes():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (CUDA IPC between two processes)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29617",
           os.path.join(ROOT, "tests", "dist_xjoin_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "xjoin parity ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_dist_operators_behind_the_abi_nccl():
    # pdrs_comm_init / pdrs_groupby_agg_dist / pdrs_join_pairs_dist on 2 or 4 processes (NCCL inside the library) against the
    # oracle on the union of the rows; with ONE visible GPU the same script runs as a single rank (world = 1: every collective
    # is a copy) so that the packing / merge / sharding code is covered on the driver's 1-GPU box too
    import torch
    n = torch.cuda.device_count()
    world = 1 if n < 2 else (2 if n < 4 else 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29619",
           os.path.join(ROOT, "tests", "dist_groupby_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "dist groupby / join parity ok" in r.stdout, r.stdout[-3000:] + r.stderr[-4000:]
    print(r.stdout[-3000:])

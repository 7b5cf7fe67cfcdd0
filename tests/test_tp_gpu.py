"""GPU, world_size >= 2 (NCCL): the tensor-parallel split of the QuantLinear stack reproduces the UNSHARDED oracle on real
GPUs (VERDICT r01 weak #3 / next #6c) — column shards, row shards + NCCL all-reduce, our peer-memory all-reduce, and the
fused matmul + all-reduce launch.  Skipped on single-GPU boxes; run with `gpurun --gpus N -- python -m pytest -m gpu`."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_tensor_parallel_stack_matches_unsharded_oracle_on_gpus():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29531",
                        os.path.join(ROOT, "tests", "tp_gpu_worker.py")], capture_output=True, text=True, timeout=400)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"tp_gpu_worker_w{world}.log"), "w") as f:  # full worker output for diagnosis
        f.write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    lines = [ln for ln in (r.stdout + r.stderr).splitlines() if "FAIL" in ln or "Error" in ln or "TP_OK" in ln]
    assert r.returncode == 0 and "TP_OK" in r.stdout, "\n".join(lines[-12:])

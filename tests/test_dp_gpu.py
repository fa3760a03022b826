"""Data-parallel correctness ON HARDWARE (needs >= 2 GPUs; skipped otherwise): the bucketed, overlapped NCCL gradient
all-reduce of dvae_b200.parallel against the mean of per-shard oracle gradients (SURVEY.md 8(e) parity definition)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("precision", ["fp16", "tf32"])
def test_two_rank_gradients_equal_mean_of_shard_oracles(precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "dp_parity_worker.py"), precision, "64"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    print(f"[dp parity {precision}] {line}")
    assert res["world"] == 2
    assert res["forward_rel_l2_max_8_non_hat"] <= 2e-3 and res["forward_rel_l2_max_hat"] <= 5e-3
    assert res["loss_rel_max"] <= 1e-3 and res["kl_rel_max"] <= 2e-3
    assert res["grad_cosine_matched"]["min"] > 0.999 and res["grad_cosine_matched"]["global"] > 0.9999
    assert res["grad_cosine_free"]["global"] > 0.99
    assert res["max_abs_difference_between_ranks"] == 0.0

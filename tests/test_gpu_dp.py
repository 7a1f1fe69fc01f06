"""Hardware parity of the batch-axis data parallelism: N NCCL ranks of the CUDA path == the single-GPU gradient of the
concatenated batch == the per-shard oracle sum (tests/dp_parity_worker.py).  Needs >= 2 GPUs (run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu`; the log of that run is profiles/r2_dp_parity_n2.txt);
skipped on a single-GPU box."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("shard", [256, 64])      # config 4 per-GPU shards at 2 and 8 GPUs
def test_nccl_ranks_equal_single_gpu_and_oracle(shard, tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    out = tmp_path / "dp.json"
    env = dict(os.environ, XG_DP_SHARD=str(shard), XG_DP_OUT=str(out), NCCL_DEBUG_FILE="/dev/null")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_parity_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    if r.returncode != 0:
        print(r.stdout[-3000:]); print(r.stderr[-12000:])
    assert r.returncode == 0, "dp_parity_worker failed (output above)"
    d = json.loads(out.read_text())
    print("dp parity:", json.dumps(d))
    assert d["eval_mode"]["worst_rel_fro"] < 1e-5 and d["train_mode"]["worst_rel_fro"] < 1e-3

"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): dopri5 with the world-scope error norm over NVLink peer
memory (gode_dopri5_fwd_world) against one process solving the whole batch."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_world_scope_norm_matches_single_process_solve():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "scripts", "world_norm_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [x for x in r.stdout.splitlines() if x.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["flags_equal"] and res["ranks_identical"], res
    assert res["dt_rel_diff"] < 1e-3 and res["sol_err"] < 2e-5 and res["param_grad_err"] < 1e-4, res

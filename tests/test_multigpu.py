"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): 2 ranks under torchrun, voxel-column shards,
NCCL all-reduce of AkA inside the library, result gathered and compared with the CPU oracle."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dist_chol,outer,lean", [("0", "1", "0"), ("1", "1", "0"), ("1", "4", "0"), ("0", "4", "0"), ("1", "4", "1")])
def test_two_rank_sharded_inversion_matches_oracle(dist_chol, outer, lean):
    """dist_chol = 1 forces the block-cyclic Cholesky with NCCL panel broadcasts (default only for M >= 8192),
    0 the replicated factorisation; outer = 4 switches both to two-level blocking (K = 512 trailing updates); lean = 1: no resident
    sensitivities, the column chunks of the refinement GEMVs are cut by the shard boundaries (int8 cases only)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + os.getpid() % 300 + 2 * int(dist_chol) + (int(outer) > 1) + 4 * int(lean)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, GEOBO_B200_DIST_CHOL=dist_chol, GEOBO_B200_CHOL_OUTER=outer, GEOBO_B200_LEAN_A=lean,
                                GEOBO_B200_LEAN_CHUNK_ROWS="4"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_OK world=2" in r.stdout

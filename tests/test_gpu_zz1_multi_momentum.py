"""The multi-rank sections of tests/mgpu_check.py beyond the pressure-correction core, when the box has >= 2 GPUs:
"momentum" = src-parallel/calcuvw.f90 through fc_calcuvw with processor faces, followed by src-parallel PISO and PIMPLE
(fc_piso) on the state it leaves; "gradients" = the `grad` dispatcher (gauss / lstsq_qr with every limiter) and calcp with
pitzDaily's own settings on the reference's shipped 2-rank decomposition.  Logs of the builder's own 2-GPU runs:
profiles/r02_mgpu_n2_*.log."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("section", ["momentum", "gradients"])
@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_two_rank_momentum_predictor(mode, section):
    """2 ranks, halo + reductions over direct NVLink stores (p2p) or NCCL collectives (nccl)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29527 + (mode == "nccl") + 2 * (section == "gradients")), os.path.join(ROOT, "tests", "mgpu_check.py")]
    env = dict(os.environ, FC_NO_P2P="0" if mode == "p2p" else "1", MGPU_SECTIONS=section)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    lines = [l for l in out.stdout.splitlines() if l.startswith("[mgpu]")]
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"mgpu_{section}_pytest_{mode}.log"), "w") as fh:
            fh.write(out.stdout + "\n--- stderr ---\n" + out.stderr)
    except OSError:
        pass
    assert any(l.startswith(f"[mgpu] {mode} ") for l in lines), "the requested communication mode did not run"
    assert out.returncode == 0 and "[mgpu] ALL OK" in lines, "\n".join(lines[-30:]) + out.stderr[-1500:]

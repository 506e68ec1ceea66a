"""GPU half of the drop-in boundary: the k-fold script's call sequence replayed through the overlay names in a fresh
process (tests/kfold_replay.py), for the plain MAE and for the scripts' default contrastive model."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("model_name", ["mae_vit_base_patch16", "contr_mae_vit_base_patch16"])
def test_kfold_call_sequence_through_overlay(model_name):
    env = dict(os.environ)
    paths = [os.path.join(ROOT, "overlay"), ROOT]
    if os.path.isdir("/root/reference/model"):           # build container only: the fall-through half of the overlay
        env["VITAE_REFERENCE_ROOT"] = "/root/reference"
        paths.append("/root/reference")
    env["PYTHONPATH"] = os.pathsep.join(paths)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "SLURM_PROCID"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "kfold_replay.py"), model_name], env=env,
                       capture_output=True, text=True, timeout=1500)
    print(r.stdout[-6000:])
    assert r.returncode == 0 and f"REPLAY_OK {model_name}" in r.stdout, r.stderr[-6000:]

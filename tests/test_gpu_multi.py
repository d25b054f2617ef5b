"""2-GPU run of the real multi-GPU path (NCCL halo exchange) against the undivided oracle result; skipped with < 2 GPUs."""
import tempfile

import pytest

pytestmark = pytest.mark.gpu


def test_two_gpu_halo_exchange_matches_undivided():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from test_halo_gloo import launch, check_against_undivided
    with tempfile.TemporaryDirectory() as d:
        launch("nccl", 2, 20, d, 29611)
        check_against_undivided(d, 2, 20)

"""Host logic of slak_b200.sparse_core.Masking on CPU tensors against the reference's golden run."""
import pytest
import torch

from _masking_replay import replay


@pytest.mark.parametrize("init", ["uniform", "ERK"])
@pytest.mark.parametrize("only_l", [False, True])
def test_masking_matches_reference_golden_cpu(init, only_l):
    mask = replay(init, only_l, torch.device("cpu"))
    assert mask.steps == 6

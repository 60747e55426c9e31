"""Host logic of slak_b200.sparse_core.Masking on CPU tensors against the reference's golden run."""
import pytest
import torch

from _masking_replay import replay


@pytest.mark.parametrize("init", ["uniform", "ERK"])
@pytest.mark.parametrize("only_l", [False, True])
def test_masking_matches_reference_golden_cpu(init, only_l):
    mask = replay(init, only_l, torch.device("cpu"))
    assert mask.steps == 6


@pytest.mark.parametrize("init,growth", [("snip", "random"), ("uniform", "gradient"), ("uniform", "momentum")])
def test_other_init_and_growth_modes_match_reference_golden_cpu(init, growth):
    """SNIP initialisation (sparse_core.py:11-47,174-181) and the gradient / momentum growth modes
    (funcs.py:196-205,227-299), pinned by runs of the reference's own code (oracle/gen_golden.py)."""
    mask = replay(init, False, torch.device("cpu"), growth_mode=growth)
    assert mask.steps == 6

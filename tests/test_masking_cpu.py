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


def test_packed_mask_state_dict_round_trip_cpu():
    """Masking.state_dict(): masks as 32-per-word bit masks + schedule state; load restores them exactly."""
    import torch
    mask = replay("uniform", True, torch.device("cpu"))
    sd = mask.state_dict()
    assert sd["steps"] == mask.steps == 6 and sd["decay"] is not None
    for n, (words, shape) in sd["masks"].items():
        assert words.dtype == torch.int32 and words.numel() == (mask.masks[n].numel() + 31) // 32
    before = {n: m.clone() for n, m in mask.masks.items()}
    for m in mask.masks.values():
        m.zero_()
    mask.load_state_dict(sd)
    for n in before:
        assert torch.equal(mask.masks[n], before[n])


def test_registries_cover_the_reference_mode_names():
    from slak_b200 import funcs
    assert set(funcs.prune_funcs) == {"magnitude", "SET", "global_magnitude"}                      # funcs.py:374-377
    assert set(funcs.growth_funcs) == {"random", "random_unfired", "momentum", "gradient", "mix", "momentum_neuron",
                                       "global_momentum_growth"}                                   # funcs.py:379-386
    assert set(funcs.redistribution_funcs) == {"momentum", "nonzero", "magnitude", "none"}         # funcs.py:388-392

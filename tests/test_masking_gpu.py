"""slak_b200.sparse_core.Masking on CUDA (fused apply kernel, radix-select prune) against the
reference's golden run: masks bit-identical, weights equal, pruned entries exact zeros."""
import numpy as np
import pytest
import torch

from _masking_replay import replay
from slak_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("init", ["uniform", "ERK"])
@pytest.mark.parametrize("only_l", [False, True])
def test_masking_matches_reference_golden_gpu(init, only_l):
    mask = replay(init, only_l, torch.device("cuda"))
    assert mask._table is not None and mask._table["count"] == len(mask.masks)   # fused path was used


@pytest.mark.parametrize("numel,k", [(1, 1), (255, 17), (4096, 0), (4096, 4096), (100003, 40001), (3_000_000, 1_234_567)])
def test_prune_kernel_equals_stable_sort(numel, k):
    g = torch.Generator().manual_seed(numel + k)
    w = torch.randn(numel, generator=g)
    w[torch.rand(numel, generator=g) < 0.3] = 0.0          # ties at zero, like masked weights
    if numel > 20:
        w[5:numel // 2] = w[5 + numel // 2 - 5: numel // 2 + numel // 2 - 5].abs().neg()   # ties at non-zero magnitudes
    mask = (torch.rand(numel, generator=g) < 0.8).float()
    want = mask.clone()
    _, idx = torch.sort(torch.abs(w), stable=True)
    want[idx[:k]] = 0.0
    lib = _lib.load()
    wd, md = w.cuda(), mask.cuda()
    ws = torch.empty(max(lib.slak_mask_prune_workspace(numel), 4096), dtype=torch.uint8, device="cuda")
    rc = lib.slak_mask_prune_magnitude(wd.data_ptr(), md.data_ptr(), numel, k, ws.data_ptr(), ws.numel(),
                                       torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "prune")
    assert torch.equal(md.cpu(), want)


def test_mask_apply_multi_tensor_ieee_semantics():
    g = torch.Generator().manual_seed(5)
    shapes = [(96, 1, 51, 5), (7,), (384, 96), (1, 1, 3, 3), (1001,)]
    ws = [torch.randn(s, generator=g) for s in shapes]
    ms = [(torch.rand(s, generator=g) < 0.6).float() for s in shapes]
    es = [torch.randn(s, generator=g) for s in shapes]
    wd, md, ed = [t.cuda() for t in ws], [t.cuda() for t in ms], [t.cuda() for t in es]
    ed[1] = None
    lib = _lib.load()
    i64 = lambda v: torch.tensor(v, dtype=torch.int64, device="cuda")
    tw, tm = i64([t.data_ptr() for t in wd]), i64([t.data_ptr() for t in md])      # tables must outlive the launch
    te, tn = i64([t.data_ptr() if t is not None else 0 for t in ed]), i64([t.numel() for t in wd])
    rc = lib.slak_mask_apply(tw.data_ptr(), tm.data_ptr(), te.data_ptr(), tn.data_ptr(), len(wd),
                             max(t.numel() for t in wd), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    _lib.check(rc, "apply")
    for w, m, e, a, b in zip(ws, ms, es, wd, ed):
        assert np.array_equal((w * m).numpy().view(np.uint32), a.cpu().numpy().view(np.uint32))   # -0.0 preserved
        if b is not None:
            assert torch.equal(e * m, b.cpu())


# ---- f4: SNIP init, gradient / momentum growth on the device kernels, packed masks -------------------------------
@pytest.mark.parametrize("init,growth", [("snip", "random"), ("uniform", "gradient"), ("uniform", "momentum")])
def test_snip_and_score_growth_modes_match_reference_golden_gpu(init, growth):
    """The golden runs of the reference's sparse_core.Masking with sparse_init=snip and growth gradient / momentum,
    replayed on CUDA: SNIP's threshold by device radix select, growth by slak_mask_grow_topk; masks bit-identical."""
    replay(init, False, torch.device("cuda"), growth_mode=growth)


@pytest.mark.parametrize("numel,k", [(1, 1), (300, 17), (4096, 0), (5000, 5000), (100003, 40001), (2_000_000, 700_001)])
def test_grow_kernel_equals_stable_descending_sort(numel, k):
    g = torch.Generator().manual_seed(3 * numel + k)
    score = torch.randn(numel, generator=g)
    score[torch.rand(numel, generator=g) < 0.2] = 0.0
    if numel > 40:
        score[7:numel // 3] = score[7 + numel // 3: 2 * (numel // 3)].abs()          # ties at non-zero magnitudes
    mask = (torch.rand(numel, generator=g) < 0.5).float()
    want = mask.clone()
    s = (score * (mask == 0).float()).abs()
    _, idx = torch.sort(s, descending=True, stable=True)
    want[idx[:k]] = 1.0
    lib = _lib.load()
    sd, md = score.cuda(), mask.cuda()
    ws = torch.empty(max(lib.slak_mask_prune_workspace(numel), 4096), dtype=torch.uint8, device="cuda")
    _lib.check(lib.slak_mask_grow_topk(sd.data_ptr(), md.data_ptr(), numel, k, ws.data_ptr(), ws.numel(),
                                       torch.cuda.current_stream().cuda_stream), "grow")
    assert torch.equal(md.cpu(), want)


@pytest.mark.parametrize("numel,k", [(1, 1), (1000, 1), (1000, 1000), (100003, 60000), (3_000_000, 1_800_000)])
def test_select_kth_largest_equals_topk(numel, k):
    x = torch.randn(numel, generator=torch.Generator().manual_seed(numel + k)).abs()
    x[::7] = 0.0
    lib = _lib.load()
    xd = x.cuda()
    ws = torch.empty(max(lib.slak_mask_prune_workspace(numel), 4096), dtype=torch.uint8, device="cuda")
    out = torch.empty(1, device="cuda")
    _lib.check(lib.slak_select_kth_largest_abs(xd.data_ptr(), numel, k, ws.data_ptr(), ws.numel(), out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream), "select")
    assert out.item() == torch.topk(x, k, sorted=True)[0][-1].item()


def test_packed_mask_checkpoint_round_trip_gpu_and_cpu_agree():
    from slak_b200.sparse_core import Masking
    mask = replay("uniform", False, torch.device("cuda"))
    sd = mask.state_dict()
    total_bits = sum(w.numel() * 32 for w, _ in sd["masks"].values())
    assert total_bits < 33 * sum(m.numel() for m in mask.masks.values()) // 32 + 32 * len(mask.masks) * 32
    for n, m in mask.masks.items():                                  # device pack == host pack, bit for bit
        assert torch.equal(Masking._pack(mask, m).cpu(), Masking._pack(mask, m.cpu()))
    before = {n: m.clone() for n, m in mask.masks.items()}
    steps, rate = mask.steps, mask.prune_rate
    for m in mask.masks.values():
        m.fill_(1.0)
    mask.steps, mask.prune_rate = 0, 0.0
    mask.load_state_dict(sd)
    assert mask.steps == steps and mask.prune_rate == rate
    for n in before:
        assert torch.equal(mask.masks[n], before[n])
    for name, p in mask._masked_params():                            # load re-applies the masks
        assert torch.all(p.data[before[name] == 0] == 0)

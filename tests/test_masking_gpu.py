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

"""INTEGRATION.md section 1 on the host side: the REFERENCE's own models/SLaK.py and sparse_core.py, imported
unmodified, pick up slak_b200/dropin/depthwise_conv2d_implicit_gemm.py in place of the CUTLASS extension module.
Runs only where the reference tree exists (the authoring container); nothing is executed on a GPU."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SLAK_REFERENCE", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not present")

SCRIPT = textwrap.dedent("""
    import sys, types
    root, ref = sys.argv[1], sys.argv[2]
    # the drop-in directory in place of the CUTLASS example directory (models/SLaK.py:9-10), then the reference itself
    sys.path[:0] = [root, root + "/slak_b200/dropin", ref]
    import importlib, os
    # oracle/ref_shims also holds a CPU stand-in of the operator module; only its timm package may be visible here
    shim = types.ModuleType("timm"); shim.__path__ = [root + "/oracle/ref_shims/timm"]
    sys.modules["timm"] = shim
    import torch
    import depthwise_conv2d_implicit_gemm as op
    assert op.__file__.startswith(root + "/slak_b200/dropin"), op.__file__
    import models.SLaK as ref_slak
    assert ref_slak.DepthWiseConv2dImplicitGEMM is op.DepthWiseConv2dImplicitGEMM
    ref_slak.use_sync_bn = False
    net = ref_slak.SLaK_tiny(kernel_size=[51, 49, 47, 13, 5], Decom=True, bn=True, width_factor=0.25)
    from slak_b200.dwconv import DepthWiseConv2dImplicitGEMM
    convs = [m for m in net.modules() if isinstance(m, DepthWiseConv2dImplicitGEMM)]
    assert len(convs) == 54 and all(isinstance(m, torch.nn.Conv2d) for m in convs)      # 18 Blocks x 3 branches
    from slak_b200 import slak
    slak.use_sync_bn = False
    ours = slak.SLaK_tiny(kernel_size=[51, 49, 47, 13, 5], Decom=True, bn=True, width_factor=0.25)
    a, b = net.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    ours.load_state_dict(a)                                 # reference checkpoints load unchanged
    # the reference's mask engine scans names / shapes of the model built on the drop-in (sparse_core.py:122-130)
    import sparse_core as ref_sparse
    torch.Tensor.cuda = lambda self, *a, **k: self
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
    args = types.SimpleNamespace(device="cpu", fix=False, update_frequency=100, only_L=True, sparse_init="uniform",
                                 sparsity=0.4, distributed=False)
    mask = ref_sparse.Masking(opt, train_loader=None, prune_rate_decay=ref_sparse.CosineDecay(0.3, 100), prune_rate=0.3,
                              prune_mode="magnitude", growth_mode="random", redistribution_mode="none", args=args)
    mask.add_module(net)
    assert len(mask.masks) == 36 and all("LoRA" in n for n in mask.masks)                # --only-L: the 36 LoRA tensors
    # without a CUDA device the operator refuses instead of computing on the CPU
    try:
        convs[0](torch.zeros(1, convs[0].in_channels, 8, 8))
    except RuntimeError as e:
        assert "CUDA" in str(e)
    else:
        raise AssertionError("CPU tensor accepted")
    print("DROPIN_OK")
""")


@pytest.mark.timeout(600)
def test_reference_model_and_mask_engine_run_on_the_dropin_module():
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, REF], capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


SCRIPT_C = textwrap.dedent("""
    import sys
    root, ref = sys.argv[1], sys.argv[2]
    ext_dir = ref + "/cutlass/examples/19_large_depthwise_conv2d_torch_extension"
    # only the native module is replaced: a directory holding just our _depthwise_conv2d_implicit_gemm_C stand-in
    import os, shutil, tempfile
    tmp = tempfile.mkdtemp()
    shutil.copy(root + "/slak_b200/dropin/_depthwise_conv2d_implicit_gemm_C.py", tmp)
    sys.path[:0] = [root, tmp, ext_dir]
    import torch
    import depthwise_conv2d_implicit_gemm as ref_op              # the REFERENCE's Python module (autograd + dispatch)
    assert ref_op.__file__.startswith(ext_dir), ref_op.__file__
    import _depthwise_conv2d_implicit_gemm_C as native
    for name in ("forward_fp32", "backward_data_fp32", "backward_filter_fp32",
                 "forward_fp16", "backward_data_fp16", "backward_filter_fp16"):           # frontend.h:3-10
        assert callable(getattr(native, name)), name
    m = ref_op.DepthWiseConv2dImplicitGEMM(8, (13, 5))
    assert isinstance(m, torch.nn.Conv2d) and tuple(m.weight.shape) == (8, 1, 13, 5)
    try:
        m(torch.zeros(1, 8, 16, 16))
    except RuntimeError as e:
        assert "CUDA" in str(e)
    else:
        raise AssertionError("CPU tensor accepted")
    print("DROPIN_C_OK")
""")


@pytest.mark.timeout(300)
def test_reference_python_module_runs_on_the_native_standin():
    ext = os.path.join(REF, "cutlass", "examples", "19_large_depthwise_conv2d_torch_extension")
    if not os.path.isdir(ext):
        pytest.skip("reference extension directory not present")
    r = subprocess.run([sys.executable, "-c", SCRIPT_C, ROOT, REF], capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert r.returncode == 0 and "DROPIN_C_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


SCRIPT_MERGE = textwrap.dedent("""
    import sys, types
    root, ref = sys.argv[1], sys.argv[2]
    sys.path[:0] = [root, root + "/slak_b200/dropin", ref]
    shim = types.ModuleType("timm"); shim.__path__ = [root + "/oracle/ref_shims/timm"]
    sys.modules["timm"] = shim
    import torch
    import models.SLaK as ref_slak
    from slak_b200 import slak
    ref_slak.use_sync_bn = False
    slak.use_sync_bn = False
    torch.manual_seed(5)
    for K, small in ((13, 5), (7, 3), (9, None)):
        kw = dict(in_channels=6, out_channels=6, kernel_size=K, stride=1, groups=6, small_kernel=small,
                  small_kernel_merged=False, Decom=False, bn=True)
        ours = slak.ReparamLargeKernelConv(**kw)
        for b in ours.branches():                     # non-trivial eval statistics and affine parameters
            b.bn.running_mean.normal_(); b.bn.running_var.uniform_(0.5, 2.0)
            b.bn.weight.data.normal_(1, 0.2); b.bn.bias.data.normal_()
            b.conv.weight.data.normal_(0, 0.1)
        theirs = ref_slak.ReparamLargeKernelConv(**kw)
        assert list(theirs.state_dict().keys()) == list(ours.state_dict().keys())
        theirs.load_state_dict(ours.state_dict())
        k1, b1 = ours.get_equivalent_kernel_bias()
        k2, b2 = theirs.get_equivalent_kernel_bias()          # models/SLaK.py:102-109 with fuse_bn :49-58
        assert torch.equal(k1, k2) and torch.equal(b1, b2), (K, small)
        ours.merge_kernel(); theirs.merge_kernel()            # models/SLaK.py:111-122
        assert list(theirs.state_dict().keys()) == list(ours.state_dict().keys()) == ["lkb_reparam.weight", "lkb_reparam.bias"]
        for k in ours.state_dict():
            assert torch.equal(ours.state_dict()[k], theirs.state_dict()[k]), (K, small, k)
    print("MERGE_OK")
""")


@pytest.mark.timeout(300)
def test_kernel_merge_and_bn_folding_equal_the_reference():
    r = subprocess.run([sys.executable, "-c", SCRIPT_MERGE, ROOT, REF], capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert r.returncode == 0 and "MERGE_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])

"""Single-task segmentation siblings sharing the multi-task kernels: nnUNet2021, BTSUNet.
CPU: oracle restatement against the fixture generated from the reference's own files, identical state_dict between the
drop-in and the oracle.  GPU: forward / loss parity, gradients reach every parameter."""
import hashlib
import os

import pytest
import torch

from oracle import torch_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"nnunet2021": lambda m: m.nnUNet2021(1, 1), "btsunet_ds": lambda m: m.BTSUNet(1, 1, 32, True),
         "btsunet": lambda m: m.BTSUNet(1, 1, 16, False)}


def _digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode()); h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def _fixture():
    return torch.load(os.path.join(HERE, "golden", "single_task.pt"), weights_only=False)


def _loss(outs, mask, dice):
    lst = outs if isinstance(outs, list) else [outs]
    return sum(dice(o, mask) / (n + 1) for n, o in enumerate(reversed(lst)))


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_and_dropin_match_reference_fixture(name):
    from multi_task_breast_cancer_b200 import models as M
    fx = _fixture()[name]
    torch.manual_seed(fx["seed"]); ora = CASES[name](O)
    torch.manual_seed(fx["seed"]); new = CASES[name](M)
    assert _digest(ora.state_dict()) == fx["state_digest"] == _digest(new.state_dict())
    assert sum(p.numel() for p in new.parameters()) == fx["n_params"]
    img, mask, _, _ = O.synthetic_batch(fx["B"], fx["S"], fx["S"], seed=fx["seed"])
    outs = ora(img)
    lst = outs if isinstance(outs, list) else [outs]
    assert len(lst) == len(fx["outputs"]) and all(torch.allclose(a, b, atol=1e-5) for a, b in zip(lst, fx["outputs"]))
    assert abs(_loss(outs, mask, O.DiceLoss()).item() - fx["loss"]) < 1e-5


def test_segmentation_factory():
    from multi_task_breast_cancer_b200 import models as M
    assert isinstance(M.init_segmentation_model("nnUNet"), M.nnUNet2021)
    assert isinstance(M.init_segmentation_model("BTSUNet", width=16, deep_supervision=True), M.BTSUNet)
    with pytest.raises(NotImplementedError):
        M.init_segmentation_model("SwinUNETR")


@pytest.mark.gpu
def test_untileable_plane_is_a_clear_error(lib):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    from multi_task_breast_cancer_b200 import models as M
    with pytest.raises(ValueError, match="not tileable"):
        M.BTSUNet(1, 1, 16, False).cuda()(torch.zeros(1, 1, 96, 96, device="cuda"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_forward_backward_parity_on_gpu(name, lib):
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    from multi_task_breast_cancer_b200 import criterions as Cr, models as M
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fx = _fixture()[name]
    torch.manual_seed(fx["seed"]); ref = CASES[name](O).cuda()
    new = CASES[name](M)
    new.load_state_dict(ref.state_dict())
    new = new.cuda()
    S = 128 if name != "btsunet" else 64
    img, mask, _, _ = O.synthetic_batch(3, S, S, device="cuda")
    ro, no = ref(img), new(img)
    rl, nl = (ro if isinstance(ro, list) else [ro]), (no if isinstance(no, list) else [no])
    assert isinstance(no, list) == isinstance(ro, list) and len(rl) == len(nl)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    # same bf16-storage floor as the multi-task parents (7.5e-2 on the BTS full-decoder head, 2.5e-2 on nnU-Net)
    last = 2.5e-2 if name == "nnunet2021" else 7.5e-2
    for i, (a, b) in enumerate(zip(nl, rl)):
        assert a.shape == b.shape and rel(a, b) < (last if i == len(nl) - 1 else 8e-2), (i, rel(a, b))
    l_new = _loss(no, mask, Cr.init_criterion_segmentation("DICE"))
    l_ref = _loss(ro, mask, O.DiceLoss())
    assert abs(l_new.item() - l_ref.item()) < 3e-3 * l_ref.item()
    l_new.backward(); l_ref.backward()
    pr = dict(ref.named_parameters())
    for n, p in new.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.shape == p.shape, n
    for n in ["output1.weight", "output1.bias"]:
        assert rel(dict(new.named_parameters())[n].grad, pr[n].grad) < 5e-2, n

"""Bring-up diagnostics (GPU): (1) the golden whole-net step with the fused residual join / single-kernel block switched
on and off, (2) per-parameter gradient distances of the TF32 AlexNet step against the TF32-operand and fp32 oracles."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from deepipr_b200 import _lib as L
from deepipr_b200 import layers, nets
from deepipr_b200.trainer import StepRunner
from oracle import passport_oracle as po
from tests.helpers import load_golden, quiet, rel_l2, seed_all


def golden(nonneg, fused):
    lib = L.load()
    lib.pp_debug_fused(1 if fused else 0)
    gm = load_golden("resnet18_private_model")
    seed_all(0)
    pk = nets.passport_kwargs_from_config(nets.resnet18_passport_config(), "bn", "random", 0.1)
    model = quiet(nets.ResNet18, "private", 10, pk)
    for m in model.modules():
        if isinstance(m, nets.BasicUnit):
            m.input_nonneg = nonneg
            m.fuse_join = nonneg
    x, t = gm["x"], gm["t"]
    seed_all(1)
    torch.randn(8, 3, 32, 32); torch.randint(0, 10, (8,))
    model = model.cuda().train()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    runner = StepRunner(model, opt, private=True, autocast=False)
    loss, sign_loss, preds = runner.forward_backward(x.cuda(), t.cuda())
    print(f"golden nonneg={nonneg} fused={fused}: logits rel", [round(rel_l2(preds[i].float().cpu(), gm["logits"][i]), 5) for i in range(2)],
          "loss", loss.item(), "ref", gm["loss"].item())


def tf32_alexnet():
    layers.set_precision("tf32")
    seed_all(0)
    pk = nets.passport_kwargs_from_config(nets.alexnet_passport_config(), "bn", "random", 0.1)
    model = quiet(nets.AlexNetCifar, "v1", 3, 10, pk)
    x = torch.randn(16, 3, 32, 32)
    t = torch.randint(0, 10, (16,))
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, layers.PassportBlock):
                c = m.conv.in_channels
                m.set_key(torch.rand(1, c, 8, 8) * 2 - 1, torch.rand(1, c, 8, 8) * 2 - 1)
    res = {}
    for mode in ("tf32", False):
        oracle = po.mirror(model, round_bf16=mode).train()
        opt_o = torch.optim.SGD(oracle.parameters(), lr=0.0)
        po.train_step(oracle, opt_o, x, t, private=False)
        res[mode] = {k: p.grad.clone() for k, p in oracle.named_parameters() if p.grad is not None}
    model = model.cuda().train()
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    runner = StepRunner(model, opt, private=False, autocast=False)
    runner.forward_backward(x.cuda(), t.cuda())
    for k, p in model.named_parameters():
        if p.grad is not None:
            print(f"tf32 grad {k:32s} vs tf32-oracle {rel_l2(p.grad, res['tf32'][k]):.3e}   vs fp32-oracle {rel_l2(p.grad, res[False][k]):.3e}"
                  f"   (tf32-oracle vs fp32-oracle {rel_l2(res['tf32'][k], res[False][k]):.3e})")
    layers.set_precision("bf16")


if __name__ == "__main__":
    for nonneg, fused in ((True, True), (False, True), (True, False), (False, False)):
        golden(nonneg, fused)
    tf32_alexnet()

"""Shared helpers of the test-suite: golden fixture loading, block construction, error metrics."""
import contextlib
import io
import os
import random

import numpy as np
import torch

from deepipr_b200 import layers

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

BLOCK_FIXTURES = [
    "v1_bn_train", "v1_bn_eval", "v1_none_s2_norelu", "v1_bn_1x1_s2", "v1_bn_keybatch2",
    "private_bn_train_2pass", "private_bn_force", "private_gn_2pass",
    "conv_bn_train", "conv_bn_s2", "conv_none", "conv_stem", "conv_bn_eval",
]


#: fixtures whose operands were NOT pre-rounded to bf16 (the reference's plain fp32 inputs)
F32OPS_FIXTURES = [
    "v1_bn_train_f32ops", "v1_bn_s2_f32ops", "private_bn_2pass_f32ops", "private_1x1_s2_keybatch2_f32ops",
]


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def seed_all(s):
    torch.manual_seed(s)
    random.seed(s)
    np.random.seed(s)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.norm().item()
    return (a - b).norm().item() / (denom if denom > 0 else 1.0)


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def product_block_from_fixture(g):
    """Build the product block described by a golden fixture and load the fixture's state_dict into it
    (exercises the _load_from_state_dict placeholder logic on the way)."""
    cfg = g["cfg"]
    kind = cfg["kind"]
    if kind == "v1":
        m = quiet(layers.PassportBlock, cfg["i"], cfg["o"], cfg["ks"], cfg["s"], cfg["pd"], cfg["kwargs"], cfg["relu"])
    elif kind == "private":
        m = quiet(layers.PassportPrivateBlock, cfg["i"], cfg["o"], cfg["ks"], cfg["s"], cfg["pd"], cfg["kwargs"])
    else:
        m = layers.ConvBlock(cfg["i"], cfg["o"], cfg["ks"], cfg["s"], cfg["pd"], bn=cfg["kwargs"]["norm_type"],
                             relu=cfg["relu"])
    state = dict(g["state"])
    if kind != "conv":
        state["conv.weight"] = state["weight"]
    m.load_state_dict(state)
    m.train(cfg["training"])
    return m


def run_block(m, g, device, make_leaf_params=True):
    """Replay the fixture's forward passes + backward on `m` (product block or oracle mirror)."""
    cfg = g["cfg"]
    x = g["x"].to(device).clone().requires_grad_(True)
    for sl in m.modules():
        if hasattr(sl, "scale_cache") and hasattr(sl, "reset"):
            sl.reset()
    total = 0
    ys = []
    for k, ind in enumerate(cfg["ind_passes"]):
        if cfg["kind"] == "v1":
            y = m(x, cfg["force_passport"])
        elif cfg["kind"] == "private":
            y = m(x, cfg["force_passport"], ind)
        else:
            y = m(x)
        total = total + (y.float() * g["r"][k].to(device)).sum()
        ys.append(y.detach().float().cpu())
    sign_total, sign_acc = 0, 0
    for sl in m.modules():
        if hasattr(sl, "scale_cache") and hasattr(sl, "reset"):
            sign_total = sign_total + sl.loss
            sign_acc = sign_acc + sl.acc
    (total + sign_total).backward()
    grads = {k: p.grad.detach().float().cpu() for k, p in m.named_parameters() if p.grad is not None}
    return dict(y=ys, dx=x.grad.detach().float().cpu(), grads=grads,
                sign_loss=torch.as_tensor(sign_total).detach().float().cpu(),
                sign_acc=torch.as_tensor(sign_acc).detach().float().cpu())

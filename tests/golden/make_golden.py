"""Generate the golden fixtures that pin oracle/passport_oracle.py (and, through it, the CUDA path).

Runs the UNMODIFIED reference (kamwoh/DeepIPR, mounted read-only at /root/reference) on CPU in fp32 and
stores inputs, parameters, outputs and gradients of its passport blocks as small .pt files next to this
script.  It is run in the build container only (the GPU box has no /root/reference); the fixtures are
committed.  Usage:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
"""
import contextlib
import io
import os
import random
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("DEEPIPR_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from models.layers.conv2d import ConvBlock  # noqa: E402
from models.layers.passportconv2d import PassportBlock  # noqa: E402
from models.layers.passportconv2d_private import PassportPrivateBlock  # noqa: E402
from models.losses.sign_loss import SignLoss  # noqa: E402


def seed_all(s):
    torch.manual_seed(s)
    random.seed(s)
    np.random.seed(s)


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def block_case(name, kind, i, o, ks, s, pd, kwargs, N, H, relu=True, training=True, seed=0, ind_passes=(0,),
               force_passport=False, round_ops=True, key_batch=1):
    """One block, forward (+backward through sum(y*r) + sign loss), everything recorded."""
    seed_all(seed)
    if kind == 'v1':
        m = quiet(PassportBlock, i, o, ks, s, pd, kwargs, relu)
    elif kind == 'private':
        m = quiet(PassportPrivateBlock, i, o, ks, s, pd, kwargs)
    else:
        m = ConvBlock(i, o, ks, s, pd, bn=kwargs['norm_type'], relu=relu)
    x = torch.randn(N, i, H, H)
    if round_ops:  # bf16-representable operands so the CUDA path sees bit-identical inputs
        x = bf16r(x)
        with torch.no_grad():
            m.conv.weight.copy_(bf16r(m.conv.weight))
    if kind != 'conv':
        key = torch.tensor(np.random.uniform(-1, 1, (key_batch, i, H, H)), dtype=torch.float32)
        skey = torch.tensor(np.random.uniform(-1, 1, (key_batch, i, H, H)), dtype=torch.float32)
        if round_ops:
            key, skey = bf16r(key), bf16r(skey)
        if key_batch == 1:
            m.set_key(key, skey)
        else:  # the block tolerates a key batch > 1 (mean over dim 0); bypass passport_selection
            kn, sn = ('key_private', 'skey_private') if kind == 'private' else ('key', 'skey')
            m.register_buffer(kn, key)
            m.register_buffer(sn, skey)
    if kind == 'private':
        with torch.no_grad():  # make the public affine non-trivial
            m.scale.copy_(torch.rand(o) + 0.5)
            m.bias.copy_(torch.randn(o) * 0.1)
    if kind == 'conv' and getattr(m, 'bn', None) is not None and hasattr(m.bn, 'weight') and m.bn.weight is not None:
        with torch.no_grad():
            m.bn.weight.copy_(torch.rand(o) + 0.5)
            m.bn.bias.copy_(torch.randn(o) * 0.1)
    m.train(training)
    # `weight` and `conv.weight` are the same Parameter in the passport blocks: store it once
    state0 = {k: v.clone() for k, v in m.state_dict().items() if not (kind != 'conv' and k == 'conv.weight')}
    x.requires_grad_(True)
    out = {'x': x.detach().clone(), 'state': state0,
           'cfg': dict(kind=kind, i=i, o=o, ks=ks, s=s, pd=pd, kwargs=kwargs, relu=relu, training=training,
                       ind_passes=list(ind_passes), force_passport=force_passport, seed=seed)}
    for sl in m.modules():
        if isinstance(sl, SignLoss):
            sl.reset()
    total = 0
    ys, rs = [], []
    for ind in ind_passes:
        if kind == 'v1':
            y = m(x, force_passport)
        elif kind == 'private':
            y = m(x, force_passport, ind)
        else:
            y = m(x)
        r = bf16r(torch.randn_like(y))
        total = total + (y * r).sum()
        ys.append(y.detach().clone())
        rs.append(r)
    sign_total = 0
    sign_acc = 0
    for sl in m.modules():
        if isinstance(sl, SignLoss):
            sign_total = sign_total + sl.loss
            sign_acc = sign_acc + sl.acc
    (total + sign_total).backward()
    out['y'] = ys
    out['r'] = rs
    out['sign_loss'] = torch.as_tensor(sign_total).detach().clone()
    out['sign_acc'] = torch.as_tensor(sign_acc).detach().clone()
    out['dx'] = x.grad.clone()
    out['grads'] = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    out['state_after'] = {k: v.clone() for k, v in m.state_dict().items() if k.startswith('bn.')}
    if kind != 'conv':
        with torch.no_grad():
            m.eval()
            if kind == 'private':
                out['gamma'] = m.get_scale(ind=1).reshape(-1).clone()
                out['beta'] = m.get_bias(ind=1).reshape(-1).clone()
            else:
                out['gamma'] = m.get_scale(True).reshape(-1).clone()
                out['beta'] = m.get_bias(True).reshape(-1).clone()
    torch.save(out, os.path.join(HERE, name + '.pt'))
    print(f'{name}: y {tuple(ys[0].shape)} sign_loss {float(out["sign_loss"]):.6f}')


def host_logic_case():
    """Signature-string decoding, passport_selection RNG replay, state_dict key sets."""
    out = {}
    seed_all(3)
    m = quiet(PassportBlock, 16, 64, 3, 1, 1, {'norm_type': 'bn', 'key_type': 'random', 'sign_loss': 0.1,
                                              'b': 'abcdefgh'})
    out['b_string'] = m.b.clone()
    out['weight_after_init'] = m.weight.detach().clone()
    seed_all(4)
    mp = quiet(PassportPrivateBlock, 16, 32, 3, 1, 1, {'norm_type': 'bn', 'key_type': 'shuffle', 'sign_loss': 0.1})
    out['private_b'] = mp.b.clone()
    out['private_weight'] = mp.weight.detach().clone()
    cand = torch.arange(20 * 16 * 4 * 4, dtype=torch.float32).view(20, 16, 4, 4)
    random.seed(7)
    out['selection_in'] = cand
    out['selection_out'] = mp.passport_selection(cand).clone()
    cand3 = torch.arange(5 * 3 * 4 * 4, dtype=torch.float32).view(5, 3, 4, 4)
    random.seed(8)
    out['selection3_out'] = mp.passport_selection(cand3).clone()
    mp.set_key(torch.randn(1, 16, 4, 4), torch.randn(1, 16, 4, 4))
    m.set_key(torch.randn(1, 16, 4, 4), torch.randn(1, 16, 4, 4))
    out['v1_keys'] = sorted(m.state_dict().keys())
    out['private_keys'] = sorted(mp.state_dict().keys())
    out['v1_params'] = sorted(k for k, _ in m.named_parameters())
    out['private_params'] = sorted(k for k, _ in mp.named_parameters())
    m.init_scale(True)
    m.init_bias(True)
    out['v1_keys_with_scale'] = sorted(m.state_dict().keys())
    cb = ConvBlock(3, 8, 3, 1, 1, bn='bn')
    out['conv_keys'] = sorted(cb.state_dict().keys())
    cbn = ConvBlock(3, 8, 3, 1, 1, bn='none')
    out['conv_none_keys'] = sorted(cbn.state_dict().keys())
    # SignLoss stand-alone
    seed_all(5)
    b = torch.sign(torch.rand(32) - 0.5)
    sl = SignLoss(0.1, b)
    scale = torch.randn(1, 32, 1, 1) * 0.2
    scale[0, 3, 0, 0] = 0.0
    scale.requires_grad_(True)
    sl.add(scale)
    sl.loss.backward()
    out['sl_b'], out['sl_scale'] = b, scale.detach().clone()
    out['sl_loss'], out['sl_acc'], out['sl_grad'] = sl.loss.detach().clone(), sl.acc.clone(), scale.grad.clone()
    torch.save(out, os.path.join(HERE, 'host_logic.pt'))
    print('host_logic: ok')


def model_case():
    """Whole ResNet18Private (CIFAR, passport layers = layer4) built by the reference with fixed seeds: parameter
    checksums, logits for both passes, and one TrainerPrivate-style step.  Only summaries are stored."""
    import json
    from experiments.trainer_private import TesterPrivate
    from models.resnet_passport_private import ResNet18Private
    cfg = json.load(open(os.path.join(REF, 'passport_configs', 'resnet18_passport.json')))

    def kwargs_of(node):
        if isinstance(node, dict):
            return {k: kwargs_of(v) for k, v in node.items()}
        kw = {'flag': bool(node), 'norm_type': 'bn', 'key_type': 'random', 'sign_loss': 0.1}
        if isinstance(node, str):
            kw['b'] = node
        return kw

    seed_all(0)
    model = quiet(ResNet18Private, num_classes=10, passport_kwargs=kwargs_of(cfg))
    out = {'param_sums': {k: v.double().sum().item() for k, v in model.state_dict().items()},
           'param_abs_sums': {k: v.double().abs().sum().item() for k, v in model.state_dict().items()}}
    seed_all(1)
    x = torch.randn(8, 3, 32, 32)
    t = torch.randint(0, 10, (8,))
    model.train()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    opt.zero_grad()
    for m in model.modules():
        if isinstance(m, SignLoss):
            m.reset()
    loss = torch.tensor(0.)
    logits = []
    for ind in range(2):
        pred = model(x, ind=ind)
        logits.append(pred.detach().clone())
        loss = loss + F.cross_entropy(pred, t)
    sign_loss = torch.tensor(0.)
    for m in model.modules():
        if isinstance(m, SignLoss):
            sign_loss = sign_loss + m.loss
    (loss + sign_loss).backward()
    out['x'], out['t'] = x, t
    out['logits'] = logits
    out['loss'], out['sign_loss'] = loss.detach().clone(), sign_loss.detach().clone()
    out['grad_norms'] = {k: p.grad.double().norm().item() for k, p in model.named_parameters() if p.grad is not None}
    # the signature itself: gamma of every passport layer on the INITIAL weights (bit-identical on both sides) ...
    def all_gammas():
        with torch.no_grad():
            return {n: m.get_scale(ind=1).reshape(-1).clone() for n, m in model.named_modules()
                    if isinstance(m, PassportPrivateBlock)}
    out['gammas_init'] = all_gammas()
    out['b'] = {n: m.b.clone() for n, m in model.named_modules() if isinstance(m, PassportPrivateBlock)}
    opt.step()
    out['gammas_after_step'] = all_gammas()          # ... and after the reference's own SGD step
    out['param_sums_after'] = {k: v.double().sum().item() for k, v in model.state_dict().items()}
    sig = quiet(TesterPrivate(model, torch.device('cpu'), verbose=False).test_signature)
    out['signature'] = sig
    # keys generated lazily by the forward (key_type='random'): store them so the replay can be checked
    out['keys'] = {k: v.clone() for k, v in model.state_dict().items()
                   if k.endswith('key_private') and 'layer4.1.convbn_2' in k}
    # checksums of every lazily generated key (all five layers must replay, not just the stored one)
    out['key_sums'] = {k: v.double().sum().item() for k, v in model.state_dict().items() if k.endswith('key_private')}
    torch.save(out, os.path.join(HERE, 'resnet18_private_model.pt'))
    print('model: loss', float(loss), 'sign', float(sign_loss))


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='', help='comma list of fixture names to (re)generate; default: all')
    only = set(filter(None, ap.parse_args().only.split(',')))
    torch.set_num_threads(8)
    bn = {'norm_type': 'bn', 'key_type': 'random', 'sign_loss': 0.1}
    none = {'norm_type': 'none', 'key_type': 'random', 'sign_loss': 0.1}
    gn = {'norm_type': 'gn', 'key_type': 'random', 'sign_loss': 0.1}
    cases = [
        ('v1_bn_train', lambda n: block_case(n, 'v1', 64, 64, 3, 1, 1, bn, N=4, H=8)),
        ('v1_bn_eval', lambda n: block_case(n, 'v1', 64, 64, 3, 1, 1, bn, N=4, H=8, training=False, seed=11)),
        ('v1_none_s2_norelu', lambda n: block_case(n, 'v1', 64, 128, 3, 2, 1, none, N=3, H=8, relu=False, seed=12)),
        ('v1_bn_1x1_s2', lambda n: block_case(n, 'v1', 64, 128, 1, 2, 0, bn, N=4, H=8, seed=13)),
        ('v1_bn_keybatch2', lambda n: block_case(n, 'v1', 64, 64, 3, 1, 1, bn, N=2, H=4, seed=14, key_batch=2)),
        ('private_bn_train_2pass', lambda n: block_case(n, 'private', 64, 64, 3, 1, 1, bn, N=4, H=8, seed=15,
                                                        ind_passes=(0, 1))),
        ('private_bn_force', lambda n: block_case(n, 'private', 64, 64, 3, 1, 1, bn, N=2, H=4, seed=16,
                                                  ind_passes=(0,), force_passport=True)),
        ('private_gn_2pass', lambda n: block_case(n, 'private', 64, 64, 3, 1, 1, gn, N=2, H=4, seed=17,
                                                  ind_passes=(0, 1))),
        ('conv_bn_train', lambda n: block_case(n, 'conv', 64, 64, 3, 1, 1, bn, N=4, H=8, seed=18)),
        ('conv_bn_s2', lambda n: block_case(n, 'conv', 64, 128, 3, 2, 1, bn, N=4, H=8, seed=19)),
        ('conv_none', lambda n: block_case(n, 'conv', 64, 64, 3, 1, 1, none, N=2, H=8, seed=20)),
        ('conv_stem', lambda n: block_case(n, 'conv', 3, 64, 3, 1, 1, bn, N=4, H=8, seed=21)),
        ('conv_bn_eval', lambda n: block_case(n, 'conv', 64, 64, 3, 1, 1, bn, N=2, H=8, seed=22, training=False)),
        # operands NOT pre-rounded to bf16: the reference's plain fp32 inputs.  The CUDA path rounds the operands of
        # the batch convolution itself, but gamma / beta / sign(gamma) come from the fp32 master weight and keys and
        # must match these bit for bit in sign.
        ('v1_bn_train_f32ops', lambda n: block_case(n, 'v1', 64, 64, 3, 1, 1, bn, N=4, H=8, seed=31, round_ops=False)),
        ('v1_bn_s2_f32ops', lambda n: block_case(n, 'v1', 128, 256, 3, 2, 1, bn, N=4, H=8, seed=32, round_ops=False)),
        ('private_bn_2pass_f32ops', lambda n: block_case(n, 'private', 128, 128, 3, 1, 1, bn, N=4, H=4, seed=33,
                                                         ind_passes=(0, 1), round_ops=False)),
        ('private_1x1_s2_keybatch2_f32ops', lambda n: block_case(n, 'private', 64, 128, 1, 2, 0, bn, N=2, H=8, seed=34,
                                                                 ind_passes=(0, 1), round_ops=False, key_batch=2)),
        ('host_logic', lambda n: host_logic_case()),
        ('resnet18_private_model', lambda n: model_case()),
    ]
    for name, fn in cases:
        if not only or name in only:
            fn(name)

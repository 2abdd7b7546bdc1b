"""BASELINE config 2 (AlexNet V1 passport, CIFAR10, fp32): the PP_DTYPE_TF32 path — fp32 activations end to end,
contractions on tcgen05 kind::tf32 — against the fp32 operators the reference calls (train_v1.py:13-29 runs the
reference in fp32; on a GPU torch's cuDNN convolutions then use TF32).

The oracle for this configuration is po.mirror(model, round_bf16='tf32'): the reference's fp32 modules with TF32
operands in the three contractions of every convolution (SURVEY §8d "fp32 with TF32 convs (config 2)").

Tolerances:
  KERNEL_TOL 5e-5  rel-L2 of a raw contraction against the fp64 result on operands cut to TF32 (the tensor core reads the
                   upper 19 bits of every fp32 operand): what is left is fp32 accumulation order
  ACT_TOL    1e-4  block outputs / logits against the TF32-operand oracle (north_star asks 1e-3)
  GRAD_TOL   1e-3  gradients against the TF32-operand oracle (an occasional ReLU-mask flip at |y| ~ 1e-6)
  TF32_TOL   2e-3  rel-L2 of forward results against the un-rounded fp32 oracle: two operands with 10 explicit mantissa
                   bits each; gradients 5e-2 (ReLU-mask flips of the elements whose pre-activation is within the TF32
                   error of zero: a sqrt-type error, the same the reference's own TF32 run has against exact fp32)
  VEC_TOL    1e-5  gamma, beta, sign loss (fp32 master weight, fp64 accumulation: independent of the conv arithmetic)
  sign(gamma) bit-exact.
"""
import pytest
import torch
import torch.nn.functional as F

from deepipr_b200 import _lib as L
from deepipr_b200 import functional as F_
from deepipr_b200 import layers, nets
from oracle import passport_oracle as po
from tests.helpers import quiet, rel_l2, seed_all

pytestmark = pytest.mark.gpu

KERNEL_TOL, ACT_TOL, GRAD_TOL, TF32_TOL, VEC_TOL = 5e-5, 1e-4, 1e-3, 2e-3, 1e-5
TF32 = L.PP_DTYPE_TF32


def tf32_cut(t):
    """fp32 -> the value a kind::tf32 instruction sees (low 13 mantissa bits dropped)."""
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.fixture(autouse=True)
def _tf32_precision():
    prev = layers.set_precision('tf32')
    yield
    layers.set_precision(prev)


GEOMS = [  # N, C, H, O, k, s, p
    (4, 192, 8, 384, 3, 1, 1),     # AlexNet features.4 (passport)
    (4, 384, 8, 256, 3, 1, 1),     # features.5 (passport)
    (3, 256, 8, 256, 3, 1, 1),     # features.6 (passport), ragged M (192 pixels)
    (2, 64, 16, 192, 5, 1, 2),     # features.2: 25 taps, 192 = 3 x 64 output columns
    (2, 3, 32, 64, 5, 1, 2),       # features.0: small-C im2col path (K = 75 -> 96)
    (5, 32, 7, 64, 3, 1, 1),       # one 32-channel chunk per tap, 245 pixels
    (2, 64, 8, 128, 3, 2, 1),      # strided: phase-decomposed data gradient
    (16, 512, 4, 512, 3, 1, 1),    # ResNet layer4 geometry: 256-wide tiles
    (2, 64, 8, 64, 1, 1, 0),       # 1x1
    (80, 64, 16, 192, 3, 1, 1),    # enough tiles for the 192-wide conv tile (features.2 at training batch sizes)
]


def _case(N, C, H, O, k, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, C, H, H, generator=g)
    w = torch.randn(O, C, k, k, generator=g) * (2.0 / (C * k * k)) ** 0.5
    return x, w


@pytest.mark.parametrize("geom", GEOMS)
def test_tf32_contractions(geom):
    """fprop / dgrad / wgrad on kind::tf32 vs fp64 convolutions of TF32-cut operands, and vs plain fp32."""
    N, C, H, O, k, s, p = geom
    spec = F_.ConvSpec(C, O, k, k, s, p)
    x, w = _case(N, C, H, O, k)
    P, Q = spec.out_hw(H, H)
    dz = torch.randn(N, O, P, Q, generator=torch.Generator().manual_seed(1))
    prep = F_.prepare_weight(w.cuda(), spec, need_dgrad=True, dtype=TF32)
    assert prep.wf.dtype == torch.float32

    xc, wc, dzc = tf32_cut(x).double(), tf32_cut(w).double(), tf32_cut(dz).double()
    z_cut = F.conv2d(xc, wc, None, s, p)
    z = F_.conv_fwd_raw(x.cuda(), prep, spec).permute(0, 3, 1, 2)
    assert z.dtype == torch.float32
    assert rel_l2(z, F.conv2d(x, w, None, s, p)) < TF32_TOL, "fprop vs fp32"
    assert rel_l2(z, z_cut) < KERNEL_TOL, "fprop vs TF32-cut operands"

    dz_nhwc = dz.permute(0, 2, 3, 1).contiguous().cuda()
    if C % 64 == 0:
        dx = F_.conv_dgrad(dz_nhwc, prep, spec, N, H, H).permute(0, 3, 1, 2)
        dx_cut = torch.nn.grad.conv2d_input((N, C, H, H), wc, dzc, s, p)
        assert dx.dtype == torch.float32
        assert rel_l2(dx, dx_cut) < KERNEL_TOL, "dgrad vs TF32-cut operands"

    dw = F_.conv_wgrad(dz_nhwc, x.cuda(), spec, dtype=TF32)
    dw_cut = torch.nn.grad.conv2d_weight(xc, (O, C, k, k), dzc, s, p)
    assert rel_l2(dw, torch.nn.grad.conv2d_weight(x, (O, C, k, k), dz, s, p)) < TF32_TOL, "wgrad vs fp32"
    assert rel_l2(dw, dw_cut) < KERNEL_TOL, "wgrad vs TF32-cut operands"


def _run_v1_block(m, x, r, force_passport=True):
    x = x.clone().requires_grad_(True)
    if m.sign_loss is not None:
        m.sign_loss.reset()
    y = m(x, force_passport)
    total = (y.float() * r).sum()
    if m.sign_loss is not None:
        total = total + m.sign_loss.loss
    total.backward()
    return dict(y=y.detach().float().cpu(), dx=x.grad.detach().float().cpu(),
                dw=m.weight.grad.detach().float().cpu(),
                sign_loss=float(m.sign_loss.loss) if m.sign_loss is not None else 0.0)


@pytest.mark.parametrize("cfg", [(192, 384, 8, 'bn', True), (384, 256, 8, 'bn', False), (256, 256, 8, 'none', True)])
def test_tf32_passport_block_matches_fp32_oracle(cfg):
    """PassportBlock (V1) forward + backward in TF32 mode vs the oracle restatement in plain fp32 on the CPU."""
    C, O, H, norm, train = cfg
    seed_all(3)
    kw = dict(key_type='random', sign_loss=0.1, norm_type=norm, flag=True, b=torch.sign(torch.rand(O) - 0.5))
    m = quiet(layers.PassportBlock, C, O, 3, 1, 1, kw)
    m.set_key(torch.randn(1, C, H, H), torch.randn(1, C, H, H))
    x = torch.randn(6, C, H, H)
    r = torch.randn(6, O, H, H)
    oracle = po.mirror(m, round_bf16='tf32')
    oracle.train(train)
    ref = _run_v1_block(oracle, x, r)
    exact = po.mirror(m, round_bf16=False)
    exact.train(train)
    ref32 = _run_v1_block(exact, x, r)
    m = m.cuda().train(train)
    out = _run_v1_block(m, x.cuda(), r.cuda())
    assert m(x.cuda(), True).dtype == torch.float32
    assert rel_l2(out["y"], ref["y"]) < ACT_TOL and rel_l2(out["y"], ref32["y"]) < TF32_TOL
    assert rel_l2(out["dx"], ref["dx"]) < GRAD_TOL and rel_l2(out["dx"], ref32["dx"]) < 5e-2
    assert rel_l2(out["dw"], ref["dw"]) < GRAD_TOL and rel_l2(out["dw"], ref32["dw"]) < 5e-2
    assert abs(out["sign_loss"] - ref["sign_loss"]) <= VEC_TOL * max(1.0, abs(ref["sign_loss"]))
    m.eval(); oracle.eval()
    with torch.no_grad():
        gamma, gref = m.get_scale(True).reshape(-1).cpu(), oracle.get_scale(True).reshape(-1)
        beta, bref = m.get_bias(True).reshape(-1).cpu(), oracle.get_bias(True).reshape(-1)
    assert rel_l2(gamma, gref) < VEC_TOL and rel_l2(beta, bref) < VEC_TOL
    assert torch.equal(torch.sign(gamma), torch.sign(gref)), "signature bits differ"


def test_tf32_alexnet_v1_step_matches_fp32_oracle():
    """One Trainer-style step (experiments/trainer.py:128-148) of AlexNet V1 with passports in features 4/5/6, fp32
    inputs, no autocast: loss, logits, every gradient and the signature against the fp32 oracle."""
    seed_all(0)
    pk = nets.passport_kwargs_from_config(nets.alexnet_passport_config(), "bn", "random", 0.1)
    model = quiet(nets.AlexNetCifar, 'v1', 3, 10, pk)
    x = torch.randn(16, 3, 32, 32)
    t = torch.randint(0, 10, (16,))
    with torch.no_grad():       # lazily created random keys: create them once, before the mirror is taken
        for m in model.modules():
            if isinstance(m, layers.PassportBlock):
                m.set_key(torch.rand(1, m.conv.in_channels, 8, 8) * 2 - 1, torch.rand(1, m.conv.in_channels, 8, 8) * 2 - 1)
    oracle = po.mirror(model, round_bf16='tf32').train()
    exact = po.mirror(model, round_bf16=False).train()
    opt_o = torch.optim.SGD(oracle.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    pred_o = oracle(x)
    ref = po.train_step(oracle, opt_o, x, t, private=False)
    po.train_step(exact, torch.optim.SGD(exact.parameters(), lr=0.0), x, t, private=False)

    model = model.cuda().train()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    from deepipr_b200.trainer import StepRunner, test_signature
    lib = L.load()
    lib.pp_launch_count(1)
    runner = StepRunner(model, opt, private=False, autocast=False)
    loss, sign_loss, preds = runner.forward_backward(x.cuda(), t.cuda())
    assert lib.pp_launch_count(0) > 0
    assert preds[0].dtype == torch.float32
    # logits: the 4096-term classifier sums cancel to O(0.1) values, which amplifies the per-activation distance
    # (< ACT_TOL, asserted block by block above) about tenfold — held to the north-star's 1e-3 (measured 2.9e-4)
    assert rel_l2(preds[0], pred_o) < 1e-3
    assert abs(loss.item() - ref["loss"]) < 1e-4 * abs(ref["loss"])
    assert abs(sign_loss.item() - ref["sign_loss"]) < 1e-4 * max(1.0, abs(ref["sign_loss"]))
    # Gradients of a 16-image batch through five BatchNorm / ReLU / max-pool stages are a chaotic function of the
    # forward values: the ReLU masks and pooling arg-maxes of the elements within the forward distance (3e-5) of a tie
    # flip, a sqrt-type error that every BatchNorm backward amplifies.  Measured: the reference's own TF32 arithmetic
    # (the TF32-operand oracle) is 4e-3 (features.6) ... 6.7e-2 (features.0) away from its exact-fp32 run; this path is
    # 2e-3 ... 1.9e-2 away from the TF32-operand oracle.  Asserted: (a) within 5e-3 of the TF32 oracle where one
    # BatchNorm lies between the loss and the layer, 4e-2 everywhere; (b) never farther from the exact fp32 gradients
    # than the reference's TF32 arithmetic is (x1.25 + 1e-3).
    g_tf32 = {k: p.grad for k, p in oracle.named_parameters() if p.grad is not None}
    g_fp32 = {k: p.grad for k, p in exact.named_parameters() if p.grad is not None}
    mine = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert set(mine) == set(g_tf32)
    for k, g in g_tf32.items():
        near = k.startswith(("features.6", "classifier"))
        assert rel_l2(mine[k], g) < (5e-3 if near else 4e-2), k
        assert rel_l2(mine[k], g_fp32[k]) < 1.25 * rel_l2(g, g_fp32[k]) + 1e-3, k
    opt.step()
    sig = test_signature(model)
    sig_o = po.test_signature(oracle.eval())
    for k in sig:
        assert abs(sig[k] - sig_o[k]) <= 1.0 / 256 + 1e-9, k      # weights one (slightly different) SGD step apart

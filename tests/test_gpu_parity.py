"""Parity of the CUDA path (through the C ABI) with the reference's golden vectors and with the CPU oracle.

Tolerances (DESIGN.md "Precision and parity"):
  ACT_TOL   1e-3  rel-L2 on bf16 activations: the block output, compared on the bf16 grid with the oracle's /
                  the reference's fp32 output rounded to bf16 (north_star: "within 1e-3 relative on bf16
                  activations/logits"); the unrounded distance is bounded by the bf16 rounding floor (1.7e-3)
  VEC_TOL   1e-5  gamma, beta, sign loss, BN statistics (fp32/fp64 arithmetic on both sides)
  GRAD_TOL  4e-3  gradients: dz and dx are bf16 tensors (operands of the tensor-core GEMMs), one bf16 rounding
                  each on top of fp32 accumulation
  sign(gamma) must be bit-exact.
"""
import pytest
import torch

from deepipr_b200 import _lib as L
from deepipr_b200 import functional as F_
from deepipr_b200 import layers, nets
from oracle import passport_oracle as po
from tests.helpers import (BLOCK_FIXTURES, F32OPS_FIXTURES, bf16r, load_golden, product_block_from_fixture, quiet,
                           rel_l2, run_block, seed_all)

pytestmark = pytest.mark.gpu

ACT_TOL, VEC_TOL, GRAD_TOL = 1e-3, 1e-5, 4e-3


def _grad(out, key):
    g = out["grads"].get(key)
    if g is None and key == "weight":
        g = out["grads"].get("conv.weight")
    if g is None and key == "conv.weight":
        g = out["grads"].get("weight")
    return g


@pytest.mark.parametrize("name", BLOCK_FIXTURES)
def test_block_matches_reference_golden(name):
    g = load_golden(name)
    m = product_block_from_fixture(g).cuda()
    out = run_block(m, g, "cuda")
    # GroupNorm takes the fused path too (groupnorm.cu): same tolerances as BatchNorm
    for k, y in enumerate(g["y"]):
        assert rel_l2(out["y"][k], bf16r(y)) < ACT_TOL, f"y[{k}]"
        assert rel_l2(out["y"][k], y) < 4e-3
    assert abs(float(out["sign_loss"]) - float(g["sign_loss"])) <= VEC_TOL * max(1.0, abs(float(g["sign_loss"])))
    assert abs(float(out["sign_acc"]) - float(g["sign_acc"])) < 1e-6
    assert rel_l2(out["dx"], g["dx"]) < GRAD_TOL
    for key, ref in g["grads"].items():
        mine = _grad(out, key)
        assert mine is not None, key
        assert rel_l2(mine, ref) < GRAD_TOL, key
    sd = m.state_dict()
    for key, ref in g["state_after"].items():
        if ref.dtype.is_floating_point:
            assert rel_l2(sd[key].cpu(), ref) < VEC_TOL, key
        else:
            assert torch.equal(sd[key].cpu(), ref), key
    if "gamma" in g:
        m.eval()
        with torch.no_grad():
            if g["cfg"]["kind"] == "private":
                gamma, beta = m.get_scale(ind=1).reshape(-1).cpu(), m.get_bias(ind=1).reshape(-1).cpu()
            else:
                gamma, beta = m.get_scale(True).reshape(-1).cpu(), m.get_bias(True).reshape(-1).cpu()
        assert rel_l2(gamma, g["gamma"]) < VEC_TOL and rel_l2(beta, g["beta"]) < VEC_TOL
        assert torch.equal(torch.sign(gamma), torch.sign(g["gamma"])), "signature bits differ from the reference"


@pytest.mark.parametrize("name", F32OPS_FIXTURES)
def test_block_matches_reference_golden_unrounded_operands(name):
    """The reference's plain fp32 inputs (nothing pre-rounded to bf16).  gamma / beta come from the fp32 master weight
    and the fp32 keys, so they match the reference to fp32 accuracy and sign(gamma) — the signature — bit for bit;
    the batch convolution rounds x and W to bf16, which shows in y and the gradients only."""
    g = load_golden(name)
    block = product_block_from_fixture(g)
    oracle = po.mirror(block, round_bf16=True)
    oracle.train(g["cfg"]["training"])
    ref = run_block(oracle, g, "cpu")
    m = block.cuda()
    out = run_block(m, g, "cuda")
    for k, y in enumerate(g["y"]):
        assert rel_l2(out["y"][k], bf16r(ref["y"][k])) < ACT_TOL, f"y[{k}] vs the bf16-operand oracle"
        assert rel_l2(out["y"][k], y) < 8e-3, f"y[{k}] vs the reference's fp32 run"   # 3 bf16 roundings (x, W, y)
    assert abs(float(out["sign_loss"]) - float(g["sign_loss"])) <= VEC_TOL * max(1.0, abs(float(g["sign_loss"])))
    assert abs(float(out["sign_acc"]) - float(g["sign_acc"])) < 1e-6
    # gradients: strict against the bf16-operand oracle; against the reference's un-rounded fp32 run the BatchNorm
    # backward (a difference of nearly equal terms) amplifies the 0.2-0.4 % operand rounding to a few percent —
    # the same distance the reference under torch.autocast(bf16) has from its own fp32 run
    assert rel_l2(out["dx"], ref["dx"]) < GRAD_TOL and rel_l2(out["dx"], g["dx"]) < 5e-2
    for key, gref in g["grads"].items():
        mine = _grad(out, key)
        assert mine is not None, key
        assert rel_l2(mine, _grad(ref, key)) < GRAD_TOL, key
        assert rel_l2(mine, gref) < 5e-2, key
    m.eval()
    with torch.no_grad():
        if g["cfg"]["kind"] == "private":
            gamma, beta = m.get_scale(ind=1).reshape(-1).cpu(), m.get_bias(ind=1).reshape(-1).cpu()
        else:
            gamma, beta = m.get_scale(True).reshape(-1).cpu(), m.get_bias(True).reshape(-1).cpu()
    assert rel_l2(gamma, g["gamma"]) < VEC_TOL and rel_l2(beta, g["beta"]) < VEC_TOL
    assert torch.equal(torch.sign(gamma), torch.sign(g["gamma"])), "signature bits differ from the reference"
    from deepipr_b200.trainer import test_signature, test_signature_per_layer
    holder = torch.nn.Sequential(m)
    want = (torch.sign(g["gamma"]) == g["state"]["b"]).float().mean().item()
    for fn in (test_signature, test_signature_per_layer):
        (det,) = fn(holder).values()
        assert det == want, fn.__name__


GEOMS = [  # N, C, H, O, k, s, p
    (2, 64, 8, 64, 1, 1, 0), (3, 64, 4, 64, 3, 1, 1), (9, 64, 4, 128, 3, 1, 1), (4, 192, 8, 384, 3, 1, 1),
    (16, 512, 4, 512, 3, 1, 1), (8, 256, 8, 512, 3, 2, 1), (8, 256, 8, 512, 1, 2, 0), (2, 64, 32, 64, 3, 1, 1),
    (5, 256, 7, 256, 3, 1, 1), (3, 128, 14, 256, 3, 2, 1), (1, 64, 4, 64, 3, 1, 1),
    # geometries that take the pixels-on-N kernel (<=128 output columns, 32x32 / 16x16 maps), incl. strided dgrad
    (3, 128, 16, 128, 3, 1, 1), (2, 64, 32, 128, 3, 2, 1), (2, 64, 32, 128, 1, 2, 0), (2, 128, 32, 64, 3, 1, 1),
    # 64 -> 64 channels: tap-paired weight gradient (row-shifted dz in the upper accumulator rows) at 2 / 4 / 8 image
    # rows per 64-pixel chunk, batch 1 included
    (3, 64, 16, 64, 3, 1, 1), (1, 64, 32, 64, 3, 1, 1), (5, 64, 8, 64, 3, 1, 1),
    # ImageNet layer1: 56x56 maps, pixels-on-N tiles of 4 image rows (N = 224)
    (2, 64, 56, 64, 3, 1, 1),
    # AlexNet channel counts at a batch with >= one 192-wide tile per SM (fprop 64 -> 192, dgrad 384 -> 192)
    (80, 64, 16, 192, 3, 1, 1), (160, 192, 8, 384, 3, 1, 1),
]


def _case(N, C, H, O, k, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = bf16r(torch.randn(N, C, H, H, generator=g))
    w = bf16r(torch.randn(O, C, k, k, generator=g) * (2.0 / (C * k * k)) ** 0.5)
    return x, w


@pytest.mark.parametrize("geom", GEOMS)
def test_conv_kernels_match_oracle_and_each_other(geom):
    """tcgen05 fprop / dgrad / wgrad vs the CPU operator the reference calls, and vs the SIMT kernels."""
    N, C, H, O, k, s, p = geom
    spec = F_.ConvSpec(C, O, k, k, s, p)
    x, w = _case(N, C, H, O, k)
    P, Q = spec.out_hw(H, H)
    dzc = bf16r(torch.randn(N, O, P, Q, generator=torch.Generator().manual_seed(1)))
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    z_ref = torch.nn.functional.conv2d(xr, wr, None, s, p)
    z_ref.backward(dzc)
    prep = F_.prepare_weight(w.cuda(), spec, True)
    dz = dzc.permute(0, 2, 3, 1).contiguous().cuda()
    res = {}
    for tag, algo in (("tc", L.PP_ALGO_TCGEN05), ("simt", L.PP_ALGO_SIMT)):
        z = F_.conv_fwd_raw(x.cuda(), prep, spec, z_f32=True, algo=algo).permute(0, 3, 1, 2)
        dx = F_.conv_dgrad(dz, prep, spec, N, H, H, algo=algo).float().permute(0, 3, 1, 2)
        dw = F_.conv_wgrad(dz, x.cuda(), spec, algo=algo)
        res[tag] = (z, dx, dw)
        assert rel_l2(z, z_ref) < 2e-5, tag
        assert rel_l2(dx, bf16r(xr.grad)) < ACT_TOL, tag     # dx is a bf16 tensor: compare on the bf16 grid
        assert rel_l2(dw, wr.grad) < 2e-5, tag
    assert rel_l2(res["tc"][0], res["simt"][0]) < 1e-5
    assert rel_l2(res["tc"][2], res["simt"][2]) < 1e-5


def _make_block(kind, i, o, ks, s, pd, norm, H, seed=0, relu=True):
    seed_all(seed)
    kw = {"norm_type": norm, "key_type": "random", "sign_loss": 0.1}
    if kind == "v1":
        m = quiet(layers.PassportBlock, i, o, ks, s, pd, kw, relu)
    elif kind == "private":
        m = quiet(layers.PassportPrivateBlock, i, o, ks, s, pd, kw)
    else:
        m = layers.ConvBlock(i, o, ks, s, pd, bn=norm, relu=relu)
    with torch.no_grad():
        m.conv.weight.copy_(bf16r(m.conv.weight))
        if kind == "private":
            m.scale.copy_(torch.rand(o) + 0.5)
            m.bias.copy_(torch.randn(o) * 0.1)
        if kind == "conv" and norm == "gn":        # non-trivial GroupNorm affine
            m.bn.weight.copy_(torch.rand(o) + 0.5)
            m.bn.bias.copy_(torch.randn(o) * 0.1)
    if kind != "conv":
        m.set_key(bf16r(torch.rand(1, i, H, H) * 2 - 1), bf16r(torch.rand(1, i, H, H) * 2 - 1))
    return m


def _fwd_bwd(mod, kind, x, dev, inds):
    mod.train()
    xx = x.to(dev).clone().requires_grad_(True)
    for sl in mod.modules():
        if hasattr(sl, "scale_cache"):
            sl.reset()
    tot, ys = 0, []
    gen = torch.Generator().manual_seed(5)
    for ind in inds:
        y = mod(xx, False, ind) if kind == "private" else (mod(xx) if kind == "conv" else mod(xx, False))
        r = bf16r(torch.randn(y.shape, generator=gen)).to(dev)
        tot = tot + (y.float() * r).sum()
        ys.append(y.detach().float().cpu())
    sl_tot = 0
    for sl in mod.modules():
        if hasattr(sl, "scale_cache"):
            sl_tot = sl_tot + sl.loss
    (tot + sl_tot).backward()
    grads = {k: p.grad.detach().float().cpu() for k, p in mod.named_parameters() if p.grad is not None}
    return dict(y=ys, dx=xx.grad.detach().float().cpu(), sl=float(sl_tot), grads=grads)


BLOCKS = [  # name, kind, i, o, ks, s, pd, norm, N, H   — the passport layers of the BASELINE configs
    ("resnet_layer4.0.convbnrelu_1", "private", 256, 512, 3, 2, 1, "bn", 32, 8),
    ("resnet_layer4.0.convbn_2", "private", 512, 512, 3, 1, 1, "bn", 32, 4),
    ("resnet_layer4.0.shortcut", "private", 256, 512, 1, 2, 0, "bn", 32, 8),
    ("alexnet_features.4", "v1", 192, 384, 3, 1, 1, "bn", 16, 8),
    ("alexnet_features.5", "v1", 384, 256, 3, 1, 1, "bn", 16, 8),
    ("v1_none_norelu", "v1", 256, 256, 3, 1, 1, "none", 8, 8),
    ("imagenet_layer4", "v1", 512, 512, 3, 1, 1, "bn", 4, 7),
    ("conv_layer1", "conv", 64, 64, 3, 1, 1, "bn", 4, 32),
    ("conv_stem", "conv", 3, 64, 3, 1, 1, "bn", 4, 32),
    ("alexnet_features.0_5x5_c3", "conv", 3, 64, 5, 1, 2, "bn", 4, 32),
    ("alexnet_features.2_5x5", "conv", 64, 192, 5, 1, 2, "bn", 4, 16),
    ("imagenet_stem_7x7_s2", "conv", 3, 64, 7, 2, 3, "bn", 2, 64),
    ("conv_none_bias", "conv", 64, 128, 3, 1, 1, "none", 4, 8),
    ("private_none", "private", 128, 128, 3, 1, 1, "none", 8, 8),
    # GroupNorm / InstanceNorm variants (--norm-type gn of flip_attack.py / passport_attack_2.py)
    ("private_gn_layer4", "private", 256, 512, 3, 2, 1, "gn", 8, 8),
    ("v1_gn_alexnet", "v1", 192, 384, 3, 1, 1, "gn", 6, 8),
    ("v1_in", "v1", 128, 128, 3, 1, 1, "in", 5, 8),
    ("private_in_1x1_s2", "private", 256, 512, 1, 2, 0, "in", 4, 8),
    ("conv_gn_affine_layer1", "conv", 64, 64, 3, 1, 1, "gn", 3, 32),
    ("conv_in", "conv", 64, 128, 3, 2, 1, "in", 3, 16),
    ("conv_gn_stem", "conv", 3, 64, 3, 1, 1, "gn", 2, 32),
]


@pytest.mark.parametrize("case", BLOCKS, ids=[c[0] for c in BLOCKS])
def test_block_matches_oracle(case):
    name, kind, i, o, ks, s, pd, norm, N, H = case
    m = _make_block(kind, i, o, ks, s, pd, norm, H, relu=(name != "v1_none_norelu"))
    x = bf16r(torch.randn(N, i, H, H, generator=torch.Generator().manual_seed(3)))
    oracle = po.mirror(m, round_bf16=True)
    inds = (0, 1) if kind == "private" else (0,)
    ref = _fwd_bwd(oracle, kind, x, "cpu", inds)
    got = _fwd_bwd(m.cuda(), kind, x, "cuda", inds)
    for k in range(len(inds)):
        assert rel_l2(got["y"][k], bf16r(ref["y"][k])) < ACT_TOL, f"y{k}"
    assert abs(got["sl"] - ref["sl"]) <= VEC_TOL * max(1.0, abs(ref["sl"]))
    if i != 3:
        assert rel_l2(got["dx"], ref["dx"]) < GRAD_TOL
    for key, gref in ref["grads"].items():
        gk = got["grads"].get(key, got["grads"].get("weight" if key == "conv.weight" else "conv.weight"))
        assert rel_l2(gk, gref) < GRAD_TOL, key
    if kind != "conv":
        with torch.no_grad():
            gg = m.get_scale(True, 1) if kind == "private" else m.get_scale(True)
            go = oracle.get_scale(True, 1)
        assert torch.equal(torch.sign(gg.reshape(-1).cpu()), torch.sign(go.reshape(-1)))
        assert rel_l2(gg.reshape(-1).cpu(), go.reshape(-1)) < VEC_TOL


def test_eval_mode_fused_epilogue_matches_training_path_kernels():
    """no-grad eval (affine folded into the conv epilogue, z never written) == oracle in eval mode."""
    m = _make_block("private", 256, 512, 3, 2, 1, "bn", 8)
    with torch.no_grad():
        m.bn.running_mean.copy_(torch.randn(512) * 0.1)
        m.bn.running_var.copy_(torch.rand(512) + 0.5)
    x = bf16r(torch.randn(16, 256, 8, 8, generator=torch.Generator().manual_seed(3)))
    oracle = po.mirror(m, round_bf16=True).eval()
    m = m.cuda().eval()
    with torch.no_grad():
        for ind in (0, 1):
            y = m(x.cuda(), False, ind).float().cpu()
            yr = oracle(x, False, ind)
            assert rel_l2(y, bf16r(yr)) < ACT_TOL


def test_signature_bits_bit_exact_full_size_layer():
    """sign(gamma) of a full-size passport layer equals the fp64 evaluation of the reference formula."""
    for (i, o, ks, s, pd, H) in ((512, 512, 3, 1, 1, 4), (256, 512, 3, 2, 1, 8), (256, 512, 1, 2, 0, 8)):
        m = _make_block("private", i, o, ks, s, pd, "bn", H, seed=7)
        w64 = m.weight.detach().double()
        ref = po.key_affine(w64, m.skey_private.double(), s, pd)
        m = m.cuda().eval()
        with torch.no_grad():
            gamma = m.get_scale(ind=1).reshape(-1).cpu()
        assert torch.equal(torch.sign(gamma), torch.sign(ref).float())
        assert rel_l2(gamma, ref) < 1e-6
        assert abs(float(m.sign_loss_private.acc) - (torch.sign(ref).float() == m.b.cpu()).float().mean().item()) < 1e-6


def test_full_size_properties_layer4_batch_1024():
    """BASELINE-size checks that need no CPU oracle: exact integer arithmetic, linearity, BN moments."""
    N, C, H, O = 1024, 512, 4, 512
    spec = F_.ConvSpec(C, O, 3, 3, 1, 1)
    g = torch.Generator(device="cuda").manual_seed(0)
    x1 = torch.randint(-2, 3, (N, C, H, H), generator=g, device="cuda").float()
    x2 = torch.randint(-2, 3, (N, C, H, H), generator=g, device="cuda").float()
    w = torch.randint(-1, 2, (O, C, 3, 3), generator=g, device="cuda").float()
    prep = F_.prepare_weight(w, spec, True)
    z1 = F_.conv_fwd_raw(x1, prep, spec, z_f32=True, algo=L.PP_ALGO_TCGEN05)
    z2 = F_.conv_fwd_raw(x2, prep, spec, z_f32=True, algo=L.PP_ALGO_TCGEN05)
    z12 = F_.conv_fwd_raw(x1 + x2, prep, spec, z_f32=True, algo=L.PP_ALGO_TCGEN05)
    assert torch.equal(z12, z1 + z2)                       # small integers: every product and sum is exact
    zs = F_.conv_fwd_raw(x1[:64], prep, spec, z_f32=True, algo=L.PP_ALGO_SIMT)
    assert torch.equal(z1[:64], zs)                        # tensor-core result == SIMT result, bit for bit
    # <dz, conv(x)> == <dgrad(dz), x> == <wgrad(dz, x), w>  (adjointness of the three kernels), integers -> exact
    dz = torch.randint(-1, 2, (N, H, H, O), generator=g, device="cuda").float()
    lhs = (dz.double() * z1.double()).sum()
    dx = F_.conv_dgrad(dz, prep, spec, N, H, H).double().permute(0, 3, 1, 2)
    dw = F_.conv_wgrad(dz, x1, spec).double()
    scale = dz.double().norm().item() * z1.double().norm().item()
    assert abs(lhs.item() - (dx * x1.double()).sum().item()) <= 1e-3 * scale     # dx is stored in bf16
    assert abs(lhs.item() - (dw * w.double()).sum().item()) <= 1e-6 * scale      # dw is fp32 (integers: exact sums)
    # BN(train) moments of the block output before ReLU: mean == beta, var == gamma^2
    m = _make_block("v1", C, O, 3, 1, 1, "bn", H, relu=False).cuda().train()
    xr = torch.randn(N, C, H, H, generator=g, device="cuda")
    y = m(xr).float()
    with torch.no_grad():
        gamma, beta = m.get_scale(True).reshape(-1), m.get_bias(True).reshape(-1)
    mean = y.mean(dim=(0, 2, 3))
    var = y.var(dim=(0, 2, 3), unbiased=False)
    assert (mean - beta).abs().max().item() < 2e-3 * (1 + beta.abs().max().item())
    assert rel_l2(var, gamma * gamma) < 5e-3


def _resnet(scheme="private", seed=0, num_classes=10):
    seed_all(seed)
    pk = nets.passport_kwargs_from_config(nets.resnet18_passport_config(), "bn", "random", 0.1)
    return quiet(nets.ResNet18, scheme, num_classes, pk)


def test_resnet18_private_step_matches_reference_golden_and_oracle():
    """Whole ResNet18 V2 step: logits / loss / sign loss / gradients / signature against the reference's own run
    (golden, fp32) and the bf16-operand oracle."""
    gm = load_golden("resnet18_private_model")
    model = _resnet()
    x, t = gm["x"], gm["t"]
    # random keys are created lazily by the first forward from numpy's RNG: replay the reference's seeds
    seed_all(1)
    torch.randn(8, 3, 32, 32); torch.randint(0, 10, (8,))
    model = model.cuda().train()
    from deepipr_b200.trainer import StepRunner
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    runner = StepRunner(model, opt, private=True, autocast=False)
    loss, sign_loss, preds = runner.forward_backward(x.cuda(), t.cuda())
    key_name = next(iter(gm["keys"]))
    assert torch.equal(model.state_dict()[key_name].cpu(), gm["keys"][key_name]), "lazy random key differs"
    # against the reference's fp32 run: bf16 operands in 20 layers => percent-level agreement on logits
    for ind in range(2):
        assert rel_l2(preds[ind].float().cpu(), gm["logits"][ind]) < 3e-2
    assert abs(loss.item() - gm["loss"].item()) < 2e-2 * gm["loss"].item()
    assert abs(sign_loss.item() - gm["sign_loss"].item()) < VEC_TOL * gm["sign_loss"].item()    # fp32 weights and keys
    gn = {k: p.grad.double().norm().item() for k, p in model.named_parameters()}
    for k in ("linear.weight", "layer4.1.convbn_2.weight", "layer4.0.convbnrelu_1.scale", "convbnrelu_1.conv.weight"):
        assert abs(gn[k] - gm["grad_norms"][k]) < 6e-2 * gm["grad_norms"][k], k
    # The signature on IDENTICAL weights and keys (the lazily created keys replay the reference's numpy stream, the
    # parameters its torch stream): every bit of every passport layer equals the reference's fp32 get_scale().
    from deepipr_b200.trainer import test_signature
    blocks = {n: m for n, m in model.named_modules() if getattr(m, "KIND", None) == "private"}
    assert set(blocks) == set(gm["gammas_init"])
    for k, v in gm["key_sums"].items():
        assert abs(model.state_dict()[k].double().sum().item() - v) < 1e-9 * max(1.0, abs(v)), k
    with torch.no_grad():
        for n, m in blocks.items():
            gam = m.get_scale(ind=1).reshape(-1).cpu()
            assert rel_l2(gam, gm["gammas_init"][n]) < VEC_TOL, n
            assert torch.equal(gam.sign(), gm["gammas_init"][n].sign()), f"signature bits differ at init in {n}"
    sig0 = test_signature(model)
    for n in blocks:
        assert sig0["private_" + n] == (gm["gammas_init"][n].sign() == gm["b"][n]).float().mean().item(), n
    model.train()
    opt.step()   # the golden signature was read after the reference's optimizer step
    # After one SGD step the two weight sets differ by lr x (bf16-operand gradient noise): measured 6.4e-4 on gamma at
    # most.  Every bit whose reference gamma is outside a 2e-3 band around zero must agree, and the detection rates
    # with them (on identical weights, above, every bit agrees).
    sig = test_signature(model)
    with torch.no_grad():
        for n, m in blocks.items():
            gam = m.get_scale(ind=1).reshape(-1).cpu()
            ref = gm["gammas_after_step"][n]
            assert (gam - ref).abs().max().item() < 2e-3, n
            decided = ref.abs() > 2e-3
            assert torch.equal(gam.sign()[decided], ref.sign()[decided]), f"signature bits differ after the step in {n}"
            flips = int((gam.sign() != ref.sign()).sum())
            assert flips <= int((~decided).sum()), n
            assert abs(sig["private_" + n] - gm["signature"]["private_" + n]) <= flips / ref.numel() + 1e-9, n


def test_resnet18_private_trajectory_and_signature_vs_oracle():
    """5 SGD steps on the GPU vs the same 5 steps of the bf16-operand oracle on the CPU."""
    model = _resnet(seed=2)
    seed_all(3)
    xs = [bf16r(torch.randn(16, 3, 32, 32)) for _ in range(5)]
    ts = [torch.randint(0, 10, (16,)) for _ in range(5)]
    with torch.no_grad():  # fix the lazily created keys so both sides share them
        for mod in model.modules():
            if getattr(mod, "KIND", None) == "private":
                c = mod.conv.in_channels
                h = 8 if mod.conv.stride[0] == 2 else 4
                mod.set_key(bf16r(torch.rand(1, c, h, h) * 2 - 1), bf16r(torch.rand(1, c, h, h) * 2 - 1))
    oracle = po.mirror(model, round_bf16=True).train()
    model = model.cuda().train()
    from deepipr_b200.trainer import StepRunner, test_signature
    opt_g = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    opt_o = torch.optim.SGD(oracle.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    runner = StepRunner(model, opt_g, private=True, autocast=False)
    for x, t in zip(xs, ts):
        loss, sl, _ = runner.step(x.cuda(), t.cuda())
        ref = po.train_step(oracle, opt_o, x, t, private=True)
        assert abs(loss.item() - ref["loss"]) < 3e-2 * abs(ref["loss"])
        assert abs(sl.item() - ref["sign_loss"]) < 2e-3 * abs(ref["sign_loss"])
    sig_g = test_signature(model)
    sig_o = po.test_signature(oracle.eval())
    assert sig_g.keys() == sig_o.keys()
    for k in sig_g:
        assert sig_g[k] == pytest.approx(sig_o[k], abs=0.01), k
    # After 5 optimisation steps the two weight trajectories differ at the bf16-gradient noise level, so a gamma
    # that sits within that noise of zero may legitimately land on either side; every other bit must agree.
    # (Bit-exactness on IDENTICAL weights is asserted in test_signature_bits_bit_exact_full_size_layer and
    #  test_block_matches_oracle.)
    for (ng, mg), (no, mo) in zip([(n, m) for n, m in model.named_modules() if getattr(m, "KIND", "") == "private"],
                                  [(n, m) for n, m in oracle.named_modules() if getattr(m, "KIND", "") == "private"]):
        with torch.no_grad():
            gam_g = mg.get_scale(ind=1).reshape(-1).cpu()
            gam_o = mo.get_scale(ind=1).reshape(-1)
        decided = gam_o.abs() > 0.05 * gam_o.abs().median()
        assert torch.equal(gam_g.sign()[decided], gam_o.sign()[decided]), f"signature bits differ in {ng}"
        assert (gam_g.sign() != gam_o.sign()).sum().item() <= 3, ng


def _fix_keys(model, kind, hw):
    with torch.no_grad():
        for mod in model.modules():
            if getattr(mod, "KIND", None) == kind:
                c = mod.conv.in_channels
                h = hw(mod)
                mod.set_key(torch.rand(1, c, h, h) * 2 - 1, torch.rand(1, c, h, h) * 2 - 1)     # plain fp32 keys


def _whole_net(net):
    seed_all(6)
    if net == "resnet18_private":
        pk = nets.passport_kwargs_from_config(nets.resnet18_passport_config(), "bn", "random", 0.1)
        model = quiet(nets.ResNet18, "private", 100, pk)                       # BASELINE config 3: CIFAR-100
        _fix_keys(model, "private", lambda m: 8 if m.conv.stride[0] == 2 else 4)
        return model, (0, 1), lambda mod, x, ind: mod(x, ind=ind)
    pk = nets.passport_kwargs_from_config(nets.alexnet_passport_config(), "bn", "random", 0.1)
    model = quiet(nets.AlexNetCifar, "v1", 3, 10, pk)
    _fix_keys(model, "v1", lambda m: 8)
    return model, (0,), lambda mod, x, ind: mod(x)


@pytest.mark.parametrize("net", ["resnet18_private", "alexnet_v1"])
def test_every_block_of_the_network_within_1e3_on_the_oracles_activations(net):
    """north_star: "within 1e-3 relative on bf16 activations".  Whole network, real weights, training-mode batch
    statistics: every block (20 conv / passport blocks of ResNet-18, 5 of AlexNet), every pass, is fed the activation
    the ORACLE fed its counterpart and must reproduce the oracle's bf16 output within 1e-3 — so one 1-ulp flip of a
    bf16 activation in an early layer (fp32 summation order) cannot masquerade as, or hide, an error further down."""
    model, inds, call = _whole_net(net)
    x = bf16r(torch.randn(32, 3, 32, 32))
    oracle = po.mirror(model, round_bf16=True).train()
    rec = {}

    def hook(name):
        def fn(mod, args, out):
            rec.setdefault(name, []).append((tuple(a.detach().clone() if torch.is_tensor(a) else a for a in args),
                                             out.detach().clone()))
        return fn

    names = [n for n, m in oracle.named_modules() if getattr(m, "KIND", None) in ("conv", "v1", "private")]
    handles = [oracle.get_submodule(n).register_forward_hook(hook(n)) for n in names]
    with torch.no_grad():
        for ind in inds:
            call(oracle, x, ind)
    for h in handles:
        h.remove()
    model = model.cuda().train()
    worst = 0.0
    with torch.no_grad():
        for n in names:
            block = model.get_submodule(n)
            assert len(rec[n]) == len(inds)
            for args, want in rec[n]:
                got = block(*[a.cuda() if torch.is_tensor(a) else a for a in args])
                err = rel_l2(got.float().cpu(), want)
                worst = max(worst, err)
                assert err < ACT_TOL, (net, n, err)
    assert len(names) == (20 if net == "resnet18_private" else 5) and worst > 0.0


@pytest.mark.parametrize("net", ["resnet18_private", "alexnet_v1"])
def test_whole_network_logits_at_the_noise_floor_of_the_bf16_activation_model(net):
    """End-to-end logits (production mode: autocast, bf16 activations, fp32 head) against the bf16-activation oracle.
    A flat 1e-3 is not a meaningful bar end to end: the oracle ITSELF moves by 0.4e-3 .. 3e-3 on these logits when only
    its accumulation precision changes (fp32 -> fp64), because an fp32 last-bit difference in front of a bf16 rounding
    flips that activation by a whole bf16 ulp and the flips compound over 20 layers.  So the distance to the oracle is
    held to that intrinsic ambiguity, measured here: max(1e-3, 1.5 x |oracle_fp32 - oracle_fp64|)."""
    import copy
    model, inds, call = _whole_net(net)
    x = bf16r(torch.randn(32, 3, 32, 32))
    oracle = po.mirror(model, round_bf16=True)
    oracle64 = copy.deepcopy(oracle).double()
    model = model.cuda()
    for training in (True, False):
        for m in (oracle, oracle64, model):
            m.train(training)
        with torch.no_grad():
            for ind in inds:
                want = call(oracle, x, ind)
                floor = rel_l2(call(oracle64, x.double(), ind).float(), want)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    got = call(model, x.cuda(), ind)
                assert got.dtype == torch.float32
                err = rel_l2(got.cpu(), want)
                assert err < max(ACT_TOL, 1.5 * floor), (net, training, ind, err, floor)


def test_alexnet_v1_step_vs_oracle():
    seed_all(0)
    pk = nets.passport_kwargs_from_config(nets.alexnet_passport_config(), "bn", "random", 0.1)
    model = quiet(nets.AlexNetCifar, "v1", 3, 10, pk)
    with torch.no_grad():
        for mod in model.modules():
            if getattr(mod, "KIND", None) == "v1":
                c = mod.conv.in_channels
                mod.set_key(bf16r(torch.rand(1, c, 8, 8) * 2 - 1), bf16r(torch.rand(1, c, 8, 8) * 2 - 1))
    x = bf16r(torch.randn(16, 3, 32, 32))
    t = torch.randint(0, 10, (16,))
    oracle = po.mirror(model, round_bf16=True).train()
    model = model.cuda().train()
    from deepipr_b200.trainer import StepRunner
    opt_g = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    opt_o = torch.optim.SGD(oracle.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    loss, sl, preds = StepRunner(model, opt_g, private=False, autocast=False).step(x.cuda(), t.cuda())
    ref = po.train_step(oracle, opt_o, x, t, private=False)
    assert abs(loss.item() - ref["loss"]) < 2e-2 * abs(ref["loss"])
    assert abs(sl.item() - ref["sign_loss"]) < 1e-3 * abs(ref["sign_loss"])


def test_flat_sgd_matches_torch_sgd():
    from deepipr_b200.parallel import FlatParams, FlatSGD
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in ((7, 5), (64,), (3, 3, 3, 3))]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    flat = FlatParams(ps)
    opt = FlatSGD(flat, lr=0.1, momentum=0.9, weight_decay=1e-4)
    opt_ref = torch.optim.SGD(ref, lr=0.1, momentum=0.9, weight_decay=1e-4)
    for step in range(4):
        opt.zero_grad()
        opt_ref.zero_grad()
        gs = [torch.randn_like(p) for p in ps]
        for p, r, g in zip(ps, ref, gs):
            p.grad.copy_(g)
            r.grad = g.clone()
        opt.step()
        opt_ref.step()
        for p, r in zip(ps, ref):
            assert torch.allclose(p, r, rtol=1e-6, atol=1e-7)


def test_imagenet_shaped_resnet18_v1_forward_backward():
    """BASELINE config 5 shape (224x224, 7x7/s2 stem + max-pool, 1000 classes) at a tiny batch: runs through the
    im2col stem path, 56/28/14/7 feature maps (tile tails), and matches the bf16-operand oracle."""
    seed_all(0)
    pk = nets.passport_kwargs_from_config(nets.resnet18_passport_config(), "bn", "random", 0.1)
    model = quiet(nets.ResNet18, "v1", 1000, pk)
    with torch.no_grad():
        for mod in model.modules():
            if getattr(mod, "KIND", None) == "v1":
                c = mod.conv.in_channels
                h = 14 if mod.conv.stride[0] == 2 else 7
                mod.set_key(bf16r(torch.rand(1, c, h, h) * 2 - 1), bf16r(torch.rand(1, c, h, h) * 2 - 1))
    x = bf16r(torch.randn(2, 3, 224, 224))
    t = torch.randint(0, 1000, (2,))
    oracle = po.mirror(model, round_bf16=True).train()
    model = model.cuda().train()
    from deepipr_b200.trainer import StepRunner
    opt_g = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    opt_o = torch.optim.SGD(oracle.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    loss, sl, preds = StepRunner(model, opt_g, private=False, autocast=False).step(x.cuda(), t.cuda())
    ref = po.train_step(oracle, opt_o, x, t, private=False)
    assert abs(loss.item() - ref["loss"]) < 3e-2 * abs(ref["loss"])
    assert abs(sl.item() - ref["sign_loss"]) < 2e-3 * abs(ref["sign_loss"])


def test_batch_one_and_odd_batches():
    """Edge cases: a single image (M = 16 rows in a 128-row tile) and a batch that is not a multiple of anything."""
    for N in (1, 7):
        m = _make_block("private", 512, 512, 3, 1, 1, "bn", 4)
        x = bf16r(torch.randn(N, 512, 4, 4, generator=torch.Generator().manual_seed(3)))
        oracle = po.mirror(m, round_bf16=True)
        ref = _fwd_bwd(oracle, "private", x, "cpu", (0, 1))
        got = _fwd_bwd(m.cuda(), "private", x, "cuda", (0, 1))
        for k in range(2):
            assert rel_l2(got["y"][k], bf16r(ref["y"][k])) < ACT_TOL
        assert rel_l2(got["dx"], ref["dx"]) < GRAD_TOL


def test_gradients_wrt_passport_keys_match_oracle():
    """passport_attack_3.py:232-270 re-registers key/skey as Parameters and optimises them: dL/dkey, dL/dskey."""
    for (i, o, ks, s, pd, H, Bk) in ((64, 128, 3, 1, 1, 8, 1), (64, 128, 3, 2, 1, 8, 2), (128, 64, 1, 2, 0, 8, 1)):
        m = _make_block("v1", i, o, ks, s, pd, "bn", H, seed=5)
        key = bf16r(torch.rand(Bk, i, H, H) * 2 - 1)
        skey = bf16r(torch.rand(Bk, i, H, H) * 2 - 1)
        x = bf16r(torch.randn(4, i, H, H, generator=torch.Generator().manual_seed(3)))
        res = {}
        for tag in ("ref", "gpu"):
            dev = "cpu" if tag == "ref" else "cuda"
            mod = po.mirror(m, round_bf16=True) if tag == "ref" else m.cuda()
            for name, val in (("key", key), ("skey", skey)):      # the attack's buffer -> Parameter swap
                if name in mod._buffers:
                    del mod._buffers[name]
                setattr(mod, name, torch.nn.Parameter(val.clone().to(dev)))
            mod.train()
            for sl in mod.modules():
                if hasattr(sl, "scale_cache"):
                    sl.reset()
            y = mod(x.to(dev))
            r = bf16r(torch.randn(y.shape, generator=torch.Generator().manual_seed(9))).to(dev)
            sl_tot = sum(sl.loss for sl in mod.modules() if hasattr(sl, "scale_cache"))
            ((y.float() * r).sum() + sl_tot).backward()
            res[tag] = (mod.key.grad.detach().float().cpu(), mod.skey.grad.detach().float().cpu())
        assert rel_l2(res["gpu"][0], res["ref"][0]) < GRAD_TOL, "dkey"
        assert rel_l2(res["gpu"][1], res["ref"][1]) < GRAD_TOL, "dskey"


def test_shared_trunk_matches_two_full_passes():
    """nets.ResNet18.share_trunk (reuse of the passport-free trunk between the ind=0 and ind=1 calls) must give
    the results of two full passes: loss, sign loss, gradients, BatchNorm running statistics."""
    from deepipr_b200.trainer import StepRunner
    xs = bf16r(torch.randn(16, 3, 32, 32, generator=torch.Generator().manual_seed(3)))
    ts = torch.randint(0, 10, (16,), generator=torch.Generator().manual_seed(4))
    out = []
    for share in (False, True):
        model = _resnet(seed=2)
        with torch.no_grad():
            for mod in model.modules():
                if getattr(mod, "KIND", None) == "private":
                    c = mod.conv.in_channels
                    h = 8 if mod.conv.stride[0] == 2 else 4
                    g = torch.Generator().manual_seed(c + h)
                    mod.set_key(bf16r(torch.rand(1, c, h, h, generator=g) * 2 - 1),
                                bf16r(torch.rand(1, c, h, h, generator=g) * 2 - 1))
        model = model.cuda().train()
        model.share_trunk = share
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
        runner = StepRunner(model, opt, private=True, autocast=True)
        x, t = xs.cuda(), ts.cuda()
        loss, sl, preds = runner.forward_backward(x, t)
        grads = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters()}
        stats = {k: v.detach().float().cpu() for k, v in model.state_dict().items() if "running" in k or "tracked" in k}
        out.append((loss.item(), sl.item(), [p.float().cpu() for p in preds], grads, stats))
    a, b = out
    assert abs(a[0] - b[0]) < 1e-3 * abs(a[0]) and abs(a[1] - b[1]) < 1e-6 * max(1.0, abs(a[1]))
    for pa, pb in zip(a[2], b[2]):
        assert rel_l2(pb, pa) < 1e-3
    for k in a[3]:
        # bf16 gradient tensors are summed at a different point; the difference is bf16 rounding noise that grows
        # with depth (17 conv layers between the loss and the stem): measured 1.1e-2 at the stem, <4e-3 in layer4
        assert rel_l2(b[3][k], a[3][k]) < 2.5e-2, k
    for k in a[4]:
        assert rel_l2(b[4][k], a[4][k]) < 1e-5, k


@pytest.mark.parametrize("shape", [(3, 64, 8, 16, 128, 3, 1, 1), (2, 128, 16, 8, 64, 3, 2, 1), (2, 64, 12, 20, 64, 1, 1, 0),
                                   (2, 64, 16, 32, 64, 3, 1, 1)])
def test_rectangular_feature_maps(shape):
    """H != W: fprop / dgrad / wgrad against the CPU operator (the reference only uses square maps, the kernels
    must not assume it)."""
    N, C, H, W, O, k, s, p = shape
    spec = F_.ConvSpec(C, O, k, k, s, p)
    g = torch.Generator().manual_seed(0)
    x = bf16r(torch.randn(N, C, H, W, generator=g))
    w = bf16r(torch.randn(O, C, k, k, generator=g) * (2.0 / (C * k * k)) ** 0.5)
    P, Q = spec.out_hw(H, W)
    dzc = bf16r(torch.randn(N, O, P, Q, generator=g))
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    z_ref = torch.nn.functional.conv2d(xr, wr, None, s, p)
    z_ref.backward(dzc)
    prep = F_.prepare_weight(w.cuda(), spec, True)
    dz = dzc.permute(0, 2, 3, 1).contiguous().cuda()
    z = F_.conv_fwd_raw(x.cuda(), prep, spec, z_f32=True).permute(0, 3, 1, 2)
    dx = F_.conv_dgrad(dz, prep, spec, N, H, W).float().permute(0, 3, 1, 2)
    dw = F_.conv_wgrad(dz, x.cuda(), spec)
    assert rel_l2(z, z_ref) < 2e-5
    assert rel_l2(dx, bf16r(xr.grad)) < ACT_TOL
    assert rel_l2(dw, wr.grad) < 2e-5


def test_conv_block_without_relu_and_in_eval_matches_oracle():
    for training in (True, False):
        m = _make_block("conv", 64, 128, 3, 1, 1, "bn", 8, relu=False)
        with torch.no_grad():
            m.bn.weight.copy_(torch.rand(128) + 0.5)
            m.bn.bias.copy_(torch.randn(128) * 0.1)
            m.bn.running_mean.copy_(torch.randn(128) * 0.1)
            m.bn.running_var.copy_(torch.rand(128) + 0.5)
        x = bf16r(torch.randn(6, 64, 8, 8, generator=torch.Generator().manual_seed(3)))
        oracle = po.mirror(m, round_bf16=True)
        outs = []
        for mod, dev in ((oracle, "cpu"), (m.cuda(), "cuda")):
            mod.train(training)
            xx = x.to(dev).clone().requires_grad_(True)
            y = mod(xx)
            r = bf16r(torch.randn(y.shape, generator=torch.Generator().manual_seed(5))).to(dev)
            (y.float() * r).sum().backward()
            outs.append((y.detach().float().cpu(), xx.grad.float().cpu(), mod.conv.weight.grad.float().cpu(),
                         mod.bn.weight.grad.float().cpu(), mod.bn.bias.grad.float().cpu()))
        ref, got = outs
        assert (got[0] < 0).any(), "no ReLU expected"
        assert rel_l2(got[0], bf16r(ref[0])) < ACT_TOL
        for a, b in zip(got[1:], ref[1:]):
            assert rel_l2(a, b) < GRAD_TOL


def test_group_and_instance_norm_run_in_the_library(monkeypatch):
    """GroupNorm / InstanceNorm blocks must not call the torch norm modules (groupnorm.cu owns that arithmetic),
    in training and in no-grad evaluation, on a ragged batch and a rectangular map."""
    def boom(self, x):
        raise AssertionError("torch norm module was called on the product path")

    for norm, cls in (("gn", torch.nn.GroupNorm), ("in", torch.nn.InstanceNorm2d)):
        m = _make_block("private", 64, 128, 3, 1, 1, norm, 8)
        x = bf16r(torch.randn(7, 64, 6, 10, generator=torch.Generator().manual_seed(11)))
        oracle = po.mirror(m, round_bf16=True)
        m = m.cuda()
        with monkeypatch.context() as mp:
            ref = _fwd_bwd(oracle, "private", x, "cpu", (0, 1))
            mp.setattr(cls, "forward", boom)
            L.load().pp_launch_count(1)
            got = _fwd_bwd(m, "private", x, "cuda", (0, 1))
            assert L.load().pp_launch_count(0) > 0
            m.eval()
            with torch.no_grad():
                y_eval = m(x.cuda(), False, 1).float().cpu()
        for k in range(2):
            assert rel_l2(got["y"][k], bf16r(ref["y"][k])) < ACT_TOL, (norm, k)
        assert rel_l2(got["dx"], ref["dx"]) < GRAD_TOL, norm
        for key, gref in ref["grads"].items():
            gk = got["grads"].get(key, got["grads"].get("weight" if key == "conv.weight" else "conv.weight"))
            assert rel_l2(gk, gref) < GRAD_TOL, (norm, key)
        oracle.eval()
        with torch.no_grad():
            y_ref = oracle(x, False, 1)
        assert rel_l2(y_eval, bf16r(y_ref)) < ACT_TOL, norm


def test_instance_norm_single_pixel_training_is_refused_like_torch():
    m = _make_block("v1", 64, 64, 3, 2, 1, "in", 2).cuda().train()
    with pytest.raises(ValueError, match="Expected more than 1 spatial element"):
        m(torch.randn(2, 64, 2, 2, device="cuda"))


def test_batched_signature_verification_equals_per_layer_path():
    """pp_signature_verify (all passport layers, one launch) returns exactly what the reference's per-layer loop
    (get_scale + sign + compare, trainer_private.py:37-71) returns, and its gamma bits are those of get_scale()."""
    from deepipr_b200.trainer import test_signature, test_signature_per_layer
    seed_all(4)
    kw = nets.passport_kwargs_from_config(nets.resnet18_passport_config(("layer3", "layer4"),
                                                                        signature="this is my signature"))
    model = quiet(nets.ResNet18, "private", 10, kw).cuda()
    with torch.no_grad():   # random keys of the right shape for every passport layer
        model.train()
        model(torch.randn(2, 3, 32, 32, device="cuda"), ind=1)
    batched = test_signature(model)
    looped = test_signature_per_layer(model)
    assert list(batched) == list(looped) and len(batched) == 10
    for k in looped:
        assert batched[k] == looped[k], k
    blocks = [m for m in model.modules() if isinstance(m, layers.PassportPrivateBlock)]
    entries = [(m.weight, m._pooled_keys()[0], m.b) for m in blocks]
    matched, Os, gammas = F_.signature_verify(entries, want_gamma=True)
    with torch.no_grad():
        for m, g, o, c in zip(blocks, gammas, Os, matched.tolist()):
            ref = m.get_scale(ind=1).reshape(-1)
            assert torch.equal(g, ref) and o == ref.numel()
            assert c == int((ref.sign() == m.b).sum())
    # a V1 block with a learnable scale (init_scale(True)) reports on that scale, as get_scale() does
    v1 = _make_block("v1", 64, 64, 3, 1, 1, "bn", 8).cuda()
    v1.init_scale(True)
    with torch.no_grad():
        v1.scale.copy_(-v1.b)
    holder = torch.nn.Sequential(v1)
    assert test_signature(holder) == {"public_0": 0.0} == test_signature_per_layer(holder)


@pytest.mark.parametrize("shape", [(3, 20, 20), (2, 12, 40), (1, 5, 70), (5, 32, 32)])
def test_direct_stem_kernels_ragged_shapes(shape):
    """stem_conv.cu (3 -> 64 channels, 3x3/s1/p1): fprop, fused statistics, eval epilogue and wgrad on maps that do
    not fill the 8 x 32 tile, against the operator the reference calls and against the SIMT kernels."""
    N, H, W = shape
    spec = F_.ConvSpec(3, 64, 3, 3, 1, 1)
    g = torch.Generator().manual_seed(21)
    x = bf16r(torch.randn(N, 3, H, W, generator=g))
    w = bf16r(torch.randn(64, 3, 3, 3, generator=g) * 0.3)
    dzc = bf16r(torch.randn(N, 64, H, W, generator=g))
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    z_ref = torch.nn.functional.conv2d(xr, wr, None, 1, 1)
    z_ref.backward(dzc)
    prep = F_.prepare_weight(w.cuda(), spec, True)
    dz = dzc.permute(0, 2, 3, 1).contiguous().cuda()
    z = F_.conv_fwd_raw(x.cuda(), prep, spec, z_f32=True).permute(0, 3, 1, 2)
    zb = F_.conv_fwd_raw(x.cuda(), prep, spec, z_f32=False).float().permute(0, 3, 1, 2)
    dw = F_.conv_wgrad(dz, x.cuda(), spec)
    assert rel_l2(z, z_ref) < 2e-5
    assert rel_l2(zb, bf16r(z_ref)) < ACT_TOL
    assert rel_l2(dw, wr.grad) < 2e-5
    z_simt = F_.conv_fwd_raw(x.cuda(), prep, spec, z_f32=True, algo=L.PP_ALGO_SIMT).permute(0, 3, 1, 2)
    assert rel_l2(z, z_simt) < 1e-5
    # whole block: training statistics (fused into the stem kernel) and the eval epilogue (affine + ReLU folded in)
    seed_all(1)
    m = layers.ConvBlock(3, 64, 3, 1, 1, bn="bn", relu=True)
    with torch.no_grad():
        m.conv.weight.copy_(w)
        m.bn.weight.copy_(torch.rand(64) + 0.5)
        m.bn.bias.copy_(torch.randn(64) * 0.1)
    oracle = po.mirror(m, round_bf16=True)
    ref = _fwd_bwd(oracle, "conv", x, "cpu", (0,))
    got = _fwd_bwd(m.cuda(), "conv", x, "cuda", (0,))
    assert rel_l2(got["y"][0], bf16r(ref["y"][0])) < ACT_TOL
    for key, gref in ref["grads"].items():
        assert rel_l2(got["grads"][key], gref) < GRAD_TOL, key
    assert rel_l2(m.bn.running_mean.cpu(), oracle.bn.running_mean) < VEC_TOL
    assert rel_l2(m.bn.running_var.cpu(), oracle.bn.running_var) < VEC_TOL
    m.eval(); oracle.eval()
    with torch.no_grad():
        assert rel_l2(m(x.cuda()).float().cpu(), bf16r(oracle(x))) < ACT_TOL


def test_direct_gradient_accumulation_equals_autograd_path():
    """ConvBlock gradients accumulated by the kernels straight into the flat buffer (PP_FLAG_ACC_*, FlatParams.direct)
    == the same step with autograd's AccumulateGrad doing `grad += g` (two passes of a V2 step), parameter for
    parameter; and a torch optimizer that detaches .grad from the flat views falls back to the autograd path."""
    from deepipr_b200.parallel import FlatParams, FlatSGD
    from deepipr_b200.trainer import StepRunner
    import bench
    x = bf16r(torch.randn(8, 3, 32, 32, generator=torch.Generator().manual_seed(2))).cuda()
    t = torch.randint(0, 10, (8,), generator=torch.Generator().manual_seed(3)).cuda()
    grads, states = [], []
    for direct in (True, False):
        model = bench.build_model(seed=0).cuda().train()
        flat = FlatParams(model.parameters())
        flat.direct = direct
        opt = FlatSGD(flat, lr=0.01, momentum=0.9, weight_decay=1e-4)
        runner = StepRunner(model, opt, private=True, autocast=True)
        L.load().pp_launch_count(1)
        runner.forward_backward(x, t)
        assert all(v == 0 for v in flat._direct_pending)
        grads.append(flat.flat_grad.clone())
        opt.step()
        runner.forward_backward(x, t)          # second step: zero_grad + accumulation into a used buffer
        grads.append(flat.flat_grad.clone())
        states.append({k: v.clone() for k, v in model.state_dict().items()})
    assert rel_l2(grads[0], grads[2]) < 1e-6 and rel_l2(grads[1], grads[3]) < 1e-6
    assert grads[0].abs().sum() > 0
    for k in states[0]:
        if states[0][k].dtype.is_floating_point:
            assert rel_l2(states[0][k], states[1][k]) < 1e-6, k
    # foreign optimizer: zero_grad(set_to_none=True) detaches .grad from the flat views
    model = bench.build_model(seed=0).cuda().train()
    flat = FlatParams(model.parameters())
    opt = torch.optim.SGD(model.parameters(), lr=0.01)
    StepRunner(model, opt, private=True, autocast=True).forward_backward(x, t)
    got = torch.cat([p.grad.reshape(-1) for p in flat.params])
    want = torch.cat([grads[0][o:o + p.numel()] for p, o in zip(flat.params, flat.offsets)])
    assert rel_l2(got, want) < 1e-6


def test_trainer_private_epoch_with_trigger_set_matches_the_reference_loop():
    """V3 (BASELINE config 4): TrainerPrivate.train over two minibatches with a trigger-set loader that is exhausted
    and restarted (trainer_private.py:131-146), then TrainerPrivate.test — against the same loop restated with the
    oracle (trainer_private.py:148-211, 73-105): dictionary keys, loss / sign-loss bookkeeping (mean vs sum), sign
    accuracy and the signature entries."""
    from deepipr_b200.trainer import TrainerPrivate
    import bench
    model = bench.build_model(seed=0)
    oracle = po.mirror(model, round_bf16=True).train()
    g = torch.Generator().manual_seed(5)
    data = [(bf16r(torch.randn(6, 3, 32, 32, generator=g)), torch.randint(0, 10, (6,), generator=g)) for _ in range(2)]
    wm = [(bf16r(torch.randn(2, 3, 32, 32, generator=g)), torch.randint(0, 10, (2,), generator=g))]
    model = model.cuda()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    trainer = TrainerPrivate(model, opt, None, torch.device("cuda"), autocast=False)
    res = trainer.train(0, data, wm)
    assert set(res) == {"loss", "sign_loss", "sign_acc", "acc_public", "acc_private", "time"}

    opt_o = torch.optim.SGD(oracle.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    it, ref = iter(wm), []
    for x, t in data:
        try:
            wx, wt = next(it)
        except StopIteration:
            it = iter(wm)
            wx, wt = next(it)
        ref.append(po.train_step(oracle, opt_o, torch.cat([x, wx]), torch.cat([t, wt]), private=True))
    loss_ref = sum(r["loss"] for r in ref) / len(ref)             # mean over batches (:186)
    sign_ref = sum(r["sign_loss"] for r in ref)                    # SUM over batches (:175, never divided)
    assert abs(res["loss"] - loss_ref) < 3e-2 * abs(loss_ref)
    assert abs(res["sign_loss"] - sign_ref) < 2e-3 * abs(sign_ref)
    accs = [float(m.acc) for m in po.sign_loss_modules(oracle)]
    assert len(accs) == 5 and abs(res["sign_acc"] - sum(accs) / len(accs)) < 0.02
    for key in ("acc_public", "acc_private"):
        assert 0.0 <= res[key] <= 100.0
        assert abs(res[key] - sum(r[key] for r in ref) / len(ref)) <= 12.5 + 1e-6      # at most one of 8 images

    out = trainer.test(data)
    sig = po.test_signature(oracle.eval())
    assert {"loss_public", "acc_public", "loss_private", "acc_private", "total_acc"} <= set(out)
    assert {"s_" + k for k in sig} <= set(out) and len(sig) == 5
    with torch.no_grad():
        for ind, key in enumerate(("public", "private")):
            lo = sum(torch.nn.functional.cross_entropy(oracle(x, ind=ind), t, reduction="sum").item() for x, t in data)
            assert abs(out["loss_" + key] - lo / 12) < 3e-2 * abs(lo / 12), key
    for k, v in sig.items():
        assert abs(out["s_" + k] - v) < 0.02, k


def test_trainer_v1_epoch_and_eval_match_the_reference_loop():
    """V1 (BASELINE config 2): Trainer.train / Trainer.test on AlexNet with passport layers 4/5/6
    (trainer.py:111-214) against the oracle's restatement of the same loop: keys and bookkeeping (sign loss is the
    MEAN over batches here, :150-152)."""
    from deepipr_b200.trainer import Trainer
    seed_all(0)
    pk = nets.passport_kwargs_from_config(nets.alexnet_passport_config(), "bn", "random", 0.1)
    model = quiet(nets.AlexNetCifar, "v1", 3, 10, pk)
    with torch.no_grad():
        for mod in model.modules():
            if getattr(mod, "KIND", None) == "v1":
                c = mod.conv.in_channels
                mod.set_key(bf16r(torch.rand(1, c, 8, 8) * 2 - 1), bf16r(torch.rand(1, c, 8, 8) * 2 - 1))
    oracle = po.mirror(model, round_bf16=True).train()
    g = torch.Generator().manual_seed(9)
    data = [(bf16r(torch.randn(8, 3, 32, 32, generator=g)), torch.randint(0, 10, (8,), generator=g)) for _ in range(2)]
    model = model.cuda()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    trainer = Trainer(model, opt, None, torch.device("cuda"), autocast=False)
    res = trainer.train(0, data)
    assert set(res) == {"loss", "sign_loss", "sign_acc", "acc", "time"}
    opt_o = torch.optim.SGD(oracle.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    ref = [po.train_step(oracle, opt_o, x, t, private=False) for x, t in data]
    assert abs(res["loss"] - sum(r["loss"] for r in ref) / 2) < 3e-2 * abs(sum(r["loss"] for r in ref) / 2)
    assert abs(res["sign_loss"] - sum(r["sign_loss"] for r in ref) / 2) < 2e-3 * abs(sum(r["sign_loss"] for r in ref) / 2)
    accs = [float(m.acc) for m in po.sign_loss_modules(oracle)]
    assert len(accs) == 3 and abs(res["sign_acc"] - sum(accs) / 3) < 0.02
    out = trainer.test(data)
    assert set(out) == {"loss", "acc", "time"}
    oracle.eval()
    with torch.no_grad():
        lo = sum(torch.nn.functional.cross_entropy(oracle(x), t, reduction="sum").item() for x, t in data) / 16
    assert abs(out["loss"] - lo) < 3e-2 * abs(lo)


def test_group_norm_conv_block_direct_gradient_accumulation():
    """PP_FLAG_ACC_* through the GroupNorm kernels (gn_dparam_kernel) and a conv bias (norm 'none'): two uses of the
    same blocks in one backward, gradients accumulated straight into FlatParams.flat_grad == autograd's sums."""
    from deepipr_b200.parallel import FlatParams
    x = bf16r(torch.randn(4, 64, 8, 8, generator=torch.Generator().manual_seed(4))).cuda()
    got = []
    for direct in (True, False):
        seed_all(3)
        net = torch.nn.Sequential(layers.ConvBlock(64, 64, 3, 1, 1, bn="gn"), layers.ConvBlock(64, 128, 3, 2, 1, bn="none"),
                                  layers.ConvBlock(128, 64, 1, 1, 0, bn="in")).cuda()
        flat = FlatParams(net.parameters())
        flat.direct = direct
        flat.zero_grad()
        L.load().pp_launch_count(1)
        (net(x).float().square().mean() + net(x * 0.5).float().abs().mean()).backward()
        assert all(v == 0 for v in flat._direct_pending)
        assert any(flat._direct_uses) == direct
        got.append(flat.flat_grad.clone())
    assert got[0].abs().sum() > 0 and rel_l2(got[0], got[1]) < 1e-6


@pytest.mark.parametrize("shape", [(7, 10, torch.float32), (1184, 10, torch.bfloat16), (130, 100, torch.float32),
                                   (64, 1000, torch.bfloat16)])
def test_fused_cross_entropy_top1_matches_torch(shape):
    """pp_ce_top1 == F.cross_entropy + accuracy()[0] (trainer_private.py:161-168, trainer.py:28-43) and its gradient."""
    from deepipr_b200.trainer import accuracy
    N, classes, dt = shape
    g = torch.Generator().manual_seed(N + classes)
    logits = (torch.randn(N, classes, generator=g) * 3).to(dt).cuda()
    target = torch.randint(0, classes, (N,), generator=g).cuda()
    a = logits.clone().requires_grad_(True)
    b = logits.clone().requires_grad_(True)
    loss, top1 = F_.ce_top1(a, target)
    (loss * 1.7).backward()
    ref = torch.nn.functional.cross_entropy(b.float(), target)
    (ref * 1.7).backward()
    assert abs(loss.item() - ref.item()) < 2e-6 * max(1.0, abs(ref.item()))
    assert abs(top1.item() - accuracy(logits.float(), target)[0].item()) < 1e-4
    tol = 1e-5 if dt == torch.float32 else 8e-3          # the gradient is returned in the logits' dtype
    assert rel_l2(a.grad, b.grad) < tol
    assert a.grad.dtype == dt and not top1.requires_grad


def test_cuda_graph_step_follows_the_eager_trajectory():
    """GraphedStepRunner: the whole V3 step replayed from one CUDA graph gives the parameters, BatchNorm statistics
    and metrics of the eager step sequence — including a learning-rate change between replays (device-side
    hyper-parameters) — and leaves no trace of its warm-up steps."""
    from deepipr_b200.parallel import FlatParams, FlatSGD
    from deepipr_b200.trainer import GraphedStepRunner, StepRunner
    import bench
    g = torch.Generator().manual_seed(8)
    batches = [(torch.randn(10, 3, 32, 32, generator=g).cuda(), torch.randint(0, 10, (10,), generator=g).cuda())
               for _ in range(4)]
    runs = []
    for graphed in (False, True):
        model = bench.build_model(seed=0).cuda().train()
        flat = FlatParams(model.parameters())
        opt = FlatSGD(flat, lr=0.05, momentum=0.9, weight_decay=1e-4)
        runner = StepRunner(model, opt, private=True, autocast=True)
        stepper = GraphedStepRunner(runner, *batches[0]) if graphed else runner
        metrics = []
        for i, (x, t) in enumerate(batches):
            if i == 2:
                opt.param_groups[0]["lr"] = 0.005
            stepper.step(x, t)
            metrics.append(runner.metrics.clone())
        torch.cuda.synchronize()
        runs.append((flat.flat.clone(), opt._buf.clone(), torch.stack(metrics),
                     {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "tracked" in k},
                     opt._steps))
    (p0, m0, met0, bn0, s0), (p1, m1, met1, bn1, s1) = runs
    assert s0 == s1 == 4
    assert rel_l2(p1, p0) < 1e-6 and rel_l2(m1, m0) < 1e-6
    assert torch.allclose(met1, met0, rtol=1e-5, atol=1e-6)
    for k in bn0:
        assert torch.equal(bn1[k], bn0[k]) if not bn0[k].dtype.is_floating_point else rel_l2(bn1[k], bn0[k]) < 1e-6, k


def _private_layer4(seed=7):
    m = _make_block("private", 512, 512, 3, 1, 1, "bn", 4, seed=seed)
    with torch.no_grad():
        m.set_key(torch.rand(1, 512, 4, 4) * 2 - 1, torch.rand(1, 512, 4, 4) * 2 - 1)
    return m


def test_single_kernel_passport_block_equals_the_kernel_sequence_at_full_size():
    """The one-kernel passport block (pp_passport_conv_fwd: conv + statistics + grid barrier + gamma/beta affine + ReLU
    from TMEM) against the kernel sequence it replaces (PP debug switch), at the BASELINE geometry: layer4 512->512,
    4x4 maps, per-GPU batch 1024 + 2 trigger images = 16416 rows = 258 tiles on 148 CTAs (two resident accumulators
    on 110 of them, a ragged last tile), both passes of the private block, forward and backward."""
    lib = L.load()
    N = 1026
    x = bf16r(torch.randn(N, 512, 4, 4, generator=torch.Generator().manual_seed(3)))
    res = {}
    try:
        for fused in (1, 0):
            lib.pp_debug_fused(fused)
            m = _private_layer4().cuda().train()
            lib.pp_launch_count(1)
            out = _fwd_bwd(m, "private", x, "cuda", (0, 1))
            out["launches"] = int(lib.pp_launch_count(0))
            out["rm"], out["rv"] = m.bn.running_mean.clone().cpu(), m.bn.running_var.clone().cpu()
            out["acc"] = float(m.sign_loss_private.acc)
            res[fused] = out
    finally:
        lib.pp_debug_fused(1)
    a, b = res[1], res[0]
    # passport pass: gemv + sign loss + conv + finalize + affine -> 1 kernel; public pass: conv + finalize + affine -> 1
    assert a["launches"] <= b["launches"] - 6, (a["launches"], b["launches"])
    for k in range(2):
        assert rel_l2(a["y"][k], b["y"][k]) < 2e-4, k           # same arithmetic; 1-ulp flips from the fp32 sum order
        assert (a["y"][k] != b["y"][k]).float().mean().item() < 2e-3, k
    assert abs(a["sl"] - b["sl"]) <= 1e-6 * max(1.0, abs(b["sl"])) and a["acc"] == b["acc"]
    assert rel_l2(a["dx"], b["dx"]) < 1e-3
    for key, gref in b["grads"].items():
        assert rel_l2(a["grads"][key], gref) < 1e-3, key
    assert rel_l2(a["rm"], b["rm"]) < 1e-5 and rel_l2(a["rv"], b["rv"]) < 1e-5
    # oracle-free property of the fused output: per-channel batch moments of the pre-ReLU activation are (beta, gamma^2)
    m = _make_block("v1", 512, 512, 3, 1, 1, "bn", 4, relu=False).cuda().train()
    y = m(x.cuda()).float()
    with torch.no_grad():
        gamma, beta = m.get_scale(True).reshape(-1), m.get_bias(True).reshape(-1)
    assert (y.mean(dim=(0, 2, 3)) - beta).abs().max().item() < 2e-3 * (1 + beta.abs().max().item())
    assert rel_l2(y.var(dim=(0, 2, 3), unbiased=False), gamma * gamma) < 5e-3


@pytest.mark.parametrize("geom", [(1026, 256, 8, 512, 3, 2, 1), (1026, 256, 8, 512, 1, 2, 0), (256, 512, 7, 512, 3, 1, 1),
                                  (40, 256, 8, 256, 3, 1, 1), (3, 512, 4, 512, 3, 1, 1)])
def test_single_kernel_block_other_geometries_vs_sequence(geom):
    """Strided 3x3 and 1x1 passport layers of layer4.0, the ImageNet 7x7 map, one 256-wide column block, a tiny
    batch: fused kernel == kernel sequence (forward outputs, statistics, gradients) for the V1 block."""
    lib = L.load()
    N, C, H, O, k, s, p = geom
    x = bf16r(torch.randn(N, C, H, H, generator=torch.Generator().manual_seed(5)))
    res = {}
    try:
        for fused in (1, 0):
            lib.pp_debug_fused(fused)
            m = _make_block("v1", C, O, k, s, p, "bn", H, seed=9).cuda().train()
            lib.pp_launch_count(1)
            res[fused] = _fwd_bwd(m, "v1", x, "cuda", (0,))
            res[fused]["launches"] = int(lib.pp_launch_count(0))
            res[fused]["rv"] = m.bn.running_var.clone().cpu()
    finally:
        lib.pp_debug_fused(1)
    a, b = res[1], res[0]
    assert a["launches"] < b["launches"], "the single-kernel path did not engage"
    assert rel_l2(a["y"][0], b["y"][0]) < 2e-4
    assert abs(a["sl"] - b["sl"]) <= 1e-6 * max(1.0, abs(b["sl"]))
    assert rel_l2(a["dx"], b["dx"]) < 1e-3 and rel_l2(a["rv"], b["rv"]) < 1e-5
    for key, gref in b["grads"].items():
        assert rel_l2(a["grads"][key], gref) < 1e-3, key


@pytest.mark.parametrize("geom", [(6, 64, 64, 1, 16), (5, 64, 128, 2, 16), (4, 256, 512, 2, 8)])
def test_residual_join_folded_into_the_block_equals_the_separate_pass(geom):
    """nets.BasicUnit with a post-ReLU input: relu(out + shortcut) (resnet_passport_private.py:78-85) has both summands
    >= 0, so it is a plain sum that convbn_2's affine pass computes (pp_conv_block_fwd_res) with an identity backward.
    Outputs and every gradient must equal, bit for bit, the general path (block, then the relu(a + b) kernel and its
    masked backward); dx is compared where x > 0 (at x == 0 == y the general path masks what an upstream ReLU masks
    again anyway)."""
    N, cin, planes, stride, H = geom
    seed_all(11)
    unit = nets.BasicUnit('normal', cin, planes, stride, None, 'bn').cuda().train()
    x0 = torch.relu(torch.randn(N, cin, H, H)).to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last)
    r = torch.randn(N, planes, H // stride, H // stride).to(torch.bfloat16).cuda()

    def run(fuse):
        unit.input_nonneg = True
        unit.fuse_join = fuse
        unit.zero_grad()
        for m in unit.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.reset_running_stats()
        x = x0.clone().requires_grad_(True)
        lib = L.load()
        lib.pp_launch_count(1)
        y = unit(x)
        (y.float() * r.float()).sum().backward()
        n = lib.pp_launch_count(0)
        return y.detach(), x.grad.detach(), {k: p.grad.detach().clone() for k, p in unit.named_parameters()}, n

    y_sep, dx_sep, g_sep, n_sep = run(False)
    y_fus, dx_fus, g_fus, n_fus = run(True)
    # the (off by default) residual-gradient link: conv1's dgrad epilogue adds the residual path's gradient
    # (functional.ResidualLink, pp_conv_block_bwd_dz dx_add) — the same sums again, bit for bit
    was, F_.RESIDUAL_LINK = F_.RESIDUAL_LINK, True
    try:
        y_lnk, dx_lnk, g_lnk, _ = run(True)
    finally:
        F_.RESIDUAL_LINK = was
    assert torch.equal(y_lnk, y_fus) and torch.equal(dx_lnk * (x0 > 0), dx_fus * (x0 > 0))
    for k in g_fus:
        assert torch.equal(g_lnk[k], g_fus[k]), k
    assert n_fus <= n_sep                # no relu(a + b) launch forward, none backward (the separate path may use the
                                         # single-kernel block for convbn_2 where it applies, hence not always - 2)
    assert torch.equal(y_sep, y_fus)
    mask = (x0 > 0)
    assert torch.equal(dx_sep * mask, dx_fus * mask)
    for k in g_sep:
        assert torch.equal(g_sep[k], g_fus[k]), k
    # and against the CPU oracle (bf16-operand model) like every other block
    unit.input_nonneg = True
    oracle = po.mirror(torch.nn.Sequential(unit), round_bf16=True).train()   # (a parent, so the unit itself is mirrored)
    xo = x0.float().cpu().requires_grad_(True)
    yo = oracle[0](xo)
    (yo * r.float().cpu()).sum().backward()
    assert rel_l2(y_fus, bf16r(yo)) < ACT_TOL
    assert rel_l2(dx_fus * mask, xo.grad * mask.cpu()) < GRAD_TOL


@pytest.mark.parametrize("case", [(2, 64, 12, 12, 3, 2, 1, torch.bfloat16), (3, 64, 9, 7, 3, 2, 1, torch.bfloat16),
                                  (2, 192, 16, 16, 2, 2, 0, torch.float32), (1, 8, 5, 5, 3, 1, 1, torch.float32),
                                  (4, 256, 8, 8, 2, 2, 0, torch.bfloat16)])
def test_nhwc_max_pool_equals_aten_bit_for_bit(case):
    """layers.MaxPool2d on channels_last tensors (pp_maxpool_fwd / _bwd) vs F.max_pool2d on the CPU — the operator
    the reference's nn.MaxPool2d layers call (alexnet_passport.py:37-38, ImageNet stem of resnet_passport*.py).  The
    inputs are post-ReLU (half of the elements tie at zero), so the first-maximum rule of the backward is exercised."""
    N, C, H, W, k, s, p, dt = case
    seed_all(5)
    x0 = torch.relu(torch.randn(N, C, H, W)).to(dt)
    pool = layers.MaxPool2d(k, s, p)
    xg = x0.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    lib = L.load()
    lib.pp_launch_count(1)
    y = pool(xg)
    assert lib.pp_launch_count(0) == 1 and y.is_contiguous(memory_format=torch.channels_last)
    xc = x0.float().requires_grad_(True)
    yc = torch.nn.functional.max_pool2d(xc, k, s, p)
    assert torch.equal(y.float().cpu(), yc.detach())
    r = torch.randn_like(yc).to(dt)
    y.backward(r.cuda())
    yc.backward(r.float())
    assert torch.equal(xg.grad.float().cpu(), xc.grad.to(dt).float())
    # NCHW-contiguous or CPU inputs are the stock module
    assert torch.equal(pool(x0.float()), yc.detach())

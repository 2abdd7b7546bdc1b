"""CPU-side behaviour of the module mirror: constructor RNG order, signature decoding, passport selection,
state_dict surface, error behaviour, C-ABI library surface.  No compute kernels are called here."""
import ctypes as C
import random
import re

import os
import pytest
import torch

from deepipr_b200 import _lib as L
from deepipr_b200 import layers, nets
from tests.helpers import load_golden, quiet, seed_all

KW = {'norm_type': 'bn', 'key_type': 'random', 'sign_loss': 0.1}


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.load()
    assert lib.pp_version() == L.PP_ABI_VERSION
    header = open(os.path.join(os.path.dirname(L._HERE), "include", "passport_sm100.h")).read()
    declared = set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_ctypes_mirror_matches_the_c_header(tmp_path):
    """Compile a probe against include/passport_sm100.h with gcc and compare struct layouts and constants with the
    ctypes mirror in deepipr_b200/_lib.py (what a cgo / JNI / ctypes binding of the ABI has to agree on)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    inc = os.path.join(os.path.dirname(L._HERE), "include")
    fields_desc = [f[0] for f in L.PPConvDesc._fields_]
    fields_sig = [f[0] for f in L.PPSigLayer._fields_]
    consts = ["PP_ABI_VERSION", "PP_NORM_NONE", "PP_NORM_BN_TRAIN", "PP_NORM_BN_EVAL", "PP_NORM_GN", "PP_ALGO_AUTO",
              "PP_ALGO_TCGEN05", "PP_ALGO_SIMT", "PP_WS_FWD", "PP_WS_BWD", "PP_FLAG_ACC_DW", "PP_FLAG_ACC_DGAMMA",
              "PP_FLAG_ACC_DBETA", "PP_SIG_MAX_LAYERS"]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "passport_sm100.h"', 'int main(void) {',
           'printf("sizeof PPConvDesc %zu\\n", sizeof(PPConvDesc));',
           'printf("sizeof PPSigLayer %zu\\n", sizeof(PPSigLayer));']
    src += [f'printf("PPConvDesc.{f} %zu\\n", offsetof(PPConvDesc, {f}));' for f in fields_desc]
    src += [f'printf("PPSigLayer.{f} %zu\\n", offsetof(PPSigLayer, {f}));' for f in fields_sig]
    src += [f'printf("{c} %d\\n", (int){c});' for c in consts]
    src += ['printf("PP_ENODEVICE %d\\n", (int)PP_ENODEVICE);', 'return 0; }']
    cfile = tmp_path / "probe.c"
    cfile.write_text("\n".join(src))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(cfile), "-o", str(exe)], check=True)
    got = dict(line.rsplit(" ", 1) for line in subprocess.run([str(exe)], capture_output=True, text=True,
                                                              check=True).stdout.splitlines())
    assert int(got["sizeof PPConvDesc"]) == C.sizeof(L.PPConvDesc)
    assert int(got["sizeof PPSigLayer"]) == C.sizeof(L.PPSigLayer)
    for f in fields_desc:
        assert int(got[f"PPConvDesc.{f}"]) == getattr(L.PPConvDesc, f).offset, f
    for f in fields_sig:
        assert int(got[f"PPSigLayer.{f}"]) == getattr(L.PPSigLayer, f).offset, f
    for c in consts:
        assert int(got[c]) == getattr(L, c), c
    assert int(got["PP_ENODEVICE"]) == -5


def test_library_fails_loudly_without_device():
    if torch.cuda.is_available():
        pytest.skip("device present")
    lib = L.load()
    d = L.PPConvDesc(N=2, C=64, H=8, W=8, O=64, kh=3, kw=3, stride=1, pad=1, norm=0, relu=1, z_f32=0, eps=1e-5,
                     momentum=0.1, algo=0, groups=0, flags=0)
    rc = lib.pp_conv_fwd_raw(C.byref(d), C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), None, 0, None)
    assert rc == -5 and "no CUDA device" in L.last_error()
    out = C.c_size_t(0)
    assert lib.pp_workspace_bytes(C.byref(d), L.PP_WS_BWD, C.byref(out)) == 0 and out.value > 0
    bad = L.PPConvDesc(N=0, C=64, H=8, W=8, O=64, kh=3, kw=3, stride=1, pad=1)
    assert lib.pp_workspace_bytes(C.byref(bad), 0, C.byref(out)) == -1


def test_cpu_tensors_raise_no_fallback():
    m = quiet(layers.PassportBlock, 64, 64, 3, 1, 1, KW)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 64, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        layers.ConvBlock(3, 8)(torch.randn(1, 3, 8, 8))


def test_constructor_rng_order_and_signature_string():
    h = load_golden("host_logic")
    seed_all(3)
    m = quiet(layers.PassportBlock, 16, 64, 3, 1, 1, dict(KW, b='abcdefgh'))
    assert torch.equal(m.b, h["b_string"])
    assert torch.equal(m.weight.detach(), h["weight_after_init"])
    bits = ''.join(format(ord(c), 'b').zfill(8) for c in 'abcdefgh')
    assert all((m.b[i] == (1 if bit == '1' else -1)) for i, bit in enumerate(bits))
    seed_all(4)
    mp = quiet(layers.PassportPrivateBlock, 16, 32, 3, 1, 1, dict(KW, key_type='shuffle'))
    assert torch.equal(mp.b, h["private_b"])
    assert torch.equal(mp.weight.detach(), h["private_weight"])
    with pytest.raises(Exception, match="Too much bit information"):
        quiet(layers.PassportBlock, 16, 32, 3, 1, 1, dict(KW, b='abcdefgh'))
    assert torch.equal(quiet(layers.PassportBlock, 4, 8, 3, 1, 1, dict(KW, b=-1)).b, -torch.ones(8))


def test_passport_selection_replays_reference_rng():
    h = load_golden("host_logic")
    mp = quiet(layers.PassportPrivateBlock, 16, 32, 3, 1, 1, KW)
    random.seed(7)
    assert torch.equal(mp.passport_selection(h["selection_in"]), h["selection_out"])
    cand3 = torch.arange(5 * 3 * 4 * 4, dtype=torch.float32).view(5, 3, 4, 4)
    random.seed(8)
    assert torch.equal(mp.passport_selection(cand3), h["selection3_out"])
    mp.set_key(h["selection_in"], h["selection_in"])
    assert mp.key_private.shape == (1, 16, 4, 4) and mp.skey_private.shape == (1, 16, 4, 4)


def test_state_dict_surface_matches_reference():
    h = load_golden("host_logic")
    m = quiet(layers.PassportBlock, 16, 64, 3, 1, 1, KW)
    mp = quiet(layers.PassportPrivateBlock, 16, 32, 3, 1, 1, KW)
    m.set_key(torch.randn(1, 16, 4, 4), torch.randn(1, 16, 4, 4))
    mp.set_key(torch.randn(1, 16, 4, 4), torch.randn(1, 16, 4, 4))
    assert sorted(m.state_dict().keys()) == h["v1_keys"]
    assert sorted(mp.state_dict().keys()) == h["private_keys"]
    assert sorted(k for k, _ in m.named_parameters()) == h["v1_params"]
    assert sorted(k for k, _ in mp.named_parameters()) == h["private_params"]
    m.init_scale(True)
    m.init_bias(True)
    assert sorted(m.state_dict().keys()) == h["v1_keys_with_scale"]
    assert sorted(layers.ConvBlock(3, 8, 3, 1, 1, bn='bn').state_dict().keys()) == h["conv_keys"]
    assert sorted(layers.ConvBlock(3, 8, 3, 1, 1, bn='none').state_dict().keys()) == h["conv_none_keys"]
    assert m.weight is m.conv.weight and m.sign_loss.b is m.b
    assert quiet(layers.PassportBlock, 4, 8, 3, 1, 1, dict(KW, sign_loss=0)).sign_loss is None


def test_load_state_dict_allocates_placeholders():
    g = load_golden("v1_bn_train")
    m = quiet(layers.PassportBlock, 64, 64, 3, 1, 1, KW)
    assert m.key is None and m.scale is None
    state = dict(g["state"])
    state["conv.weight"] = state["weight"]
    state["scale"] = torch.full((64,), 2.0)
    state["bias"] = torch.full((64,), -1.0)
    m.load_state_dict(state)
    assert torch.equal(m.key, state["key"]) and torch.equal(m.skey, state["skey"])
    assert isinstance(m.scale, torch.nn.Parameter) and torch.equal(m.scale.detach(), state["scale"])
    assert torch.equal(m.get_scale().reshape(-1).detach(), state["scale"])      # public path, no kernels needed
    assert torch.equal(m.get_bias().reshape(-1).detach(), state["bias"])


def test_sign_loss_protocol():
    sl = layers.SignLoss(0.1, torch.ones(4))
    assert sl.loss == 0 and sl.acc == 0 and sl.scale_cache is None
    with pytest.raises(Exception, match="scale_cache is None"):
        sl.get_loss()
    with pytest.raises(Exception, match="scale_cache is None"):
        sl.get_acc()
    sl._add_fused(torch.tensor([1., -1., .05, 0.]).view(1, 4, 1, 1), torch.tensor(0.5), torch.tensor(0.25))
    assert float(sl.loss) == 0.5 and float(sl.acc) == 0.25
    assert abs(float(sl.get_loss()) - 0.1 * (0.0 + 1.1 + 0.05 + 0.1)) < 1e-6
    assert abs(float(sl.get_acc()) - 0.5) < 1e-6
    sl.reset()
    assert sl.loss == 0 and sl.scale_cache is None
    sl.set_b(-torch.ones(4))
    assert torch.equal(sl.b, -torch.ones(4))


def test_resnet18_private_parameters_replay_reference_construction():
    """Same seeds => bit-identical parameters as the reference's ResNet18Private (RNG consumption order)."""
    gm = load_golden("resnet18_private_model")
    seed_all(0)
    pk = nets.passport_kwargs_from_config(nets.resnet18_passport_config(), 'bn', 'random', 0.1)
    model = quiet(nets.ResNet18, 'private', 10, pk)
    sd = model.state_dict()
    missing = set(gm["param_sums"]) - set(sd)
    # key_private/skey_private only exist in the golden dict after its lazy random-key forward
    assert all(k.endswith("key_private") for k in missing), missing
    for k, v in sd.items():
        assert abs(v.double().sum().item() - gm["param_sums"][k]) < 1e-9, k
        assert abs(v.double().abs().sum().item() - gm["param_abs_sums"][k]) < 1e-9, k


def test_config_helpers():
    cfg = nets.resnet18_passport_config(signature="this is my signature")
    assert cfg['layer4']['1']['convbn_2'] == "this is my signature" and cfg['layer1']['0']['convbnrelu_1'] is False
    assert 'shortcut' in cfg['layer2']['0'] and 'shortcut' not in cfg['layer1']['0']
    kw = nets.passport_kwargs_from_config(cfg, 'bn', 'shuffle', 0.1)
    leaf = kw['layer4']['1']['convbn_2']
    assert leaf['flag'] is True and leaf['b'] == "this is my signature" and leaf['key_type'] == 'shuffle'
    assert nets.alexnet_passport_config() == {'0': False, '2': False, '4': True, '5': True, '6': True}


def test_precision_switch_selects_the_arithmetic_per_call(monkeypatch):
    """layers.set_precision / a block's own `precision`: TF32 applies to fp32 inputs outside autocast on batch-norm /
    plain blocks only; everything else keeps the bf16 path (DESIGN.md 'TF32')."""
    import pytest
    from deepipr_b200 import _lib as L
    from deepipr_b200 import layers
    blk = layers.ConvBlock(32, 64, 3, 1, 1, bn='bn')
    x32, x16 = torch.zeros(1, 32, 4, 4), torch.zeros(1, 32, 4, 4, dtype=torch.bfloat16)
    prev = layers.set_precision('bf16')
    try:
        assert blk._dtype(x32, L.PP_NORM_BN_TRAIN) == L.PP_DTYPE_BF16
        assert layers.set_precision('tf32') == 'bf16' and layers.get_precision() == 'tf32'
        assert blk._dtype(x32, L.PP_NORM_BN_TRAIN) == L.PP_DTYPE_TF32
        assert blk._dtype(x16, L.PP_NORM_BN_TRAIN) == L.PP_DTYPE_BF16            # bf16 input
        assert blk._dtype(x32, L.PP_NORM_GN) == L.PP_DTYPE_BF16                  # group / instance norm: bf16 kernels
        with monkeypatch.context() as mp:                                        # inside torch.autocast('cuda', bf16)
            mp.setattr(torch, "is_autocast_enabled", lambda *a: True)
            assert blk._dtype(x32, L.PP_NORM_BN_TRAIN) == L.PP_DTYPE_BF16
        blk.precision = 'bf16'                                                   # per-block override
        assert blk._dtype(x32, L.PP_NORM_BN_TRAIN) == L.PP_DTYPE_BF16
        with pytest.raises(ValueError):
            layers.set_precision('fp8')
    finally:
        layers.set_precision(prev)


def test_residual_join_is_folded_only_where_both_summands_are_nonnegative():
    """nets.BasicUnit: relu(out + shortcut) is a plain sum (folded into convbn_2 / identity backward) iff convbn_2 ends
    in a ReLU and the shortcut is a ReLU block's output or a post-ReLU unit input (set by ResNet18 for its units)."""
    from deepipr_b200 import nets
    ident = nets.BasicUnit('normal', 64, 64, 1, None, 'bn')
    proj = nets.BasicUnit('normal', 64, 128, 2, None, 'bn')
    assert not ident._join_is_plain_sum()          # stand-alone unit: nothing known about the sign of its input
    ident.input_nonneg = True
    assert ident._join_is_plain_sum()
    assert proj._join_is_plain_sum()               # projection shortcut ends in a ReLU like every reference block
    proj.fuse_join = False
    assert not proj._join_is_plain_sum()
    proj.fuse_join = True
    proj.shortcut.relu = None
    assert not proj._join_is_plain_sum()
    ident.convbn_2.relu = None
    assert not ident._join_is_plain_sum()
    net = quiet(nets.ResNet18, 'private', 10, nets.passport_kwargs_from_config(nets.resnet18_passport_config(), 'bn',
                                                                                'random', 0.1))
    assert all(u.input_nonneg for u in net._units())


def test_max_pool_module_is_the_stock_module_off_the_gpu():
    from deepipr_b200 import layers
    pool = layers.MaxPool2d(3, 2, 1)
    x = torch.randn(2, 8, 9, 9)
    assert torch.equal(pool(x), torch.nn.functional.max_pool2d(x, 3, 2, 1))
    assert list(pool.state_dict().keys()) == [] and isinstance(pool, torch.nn.MaxPool2d)

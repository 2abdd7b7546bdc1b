"""The unmodified reference as a test partner.

CPU part (runs everywhere): the reference archive (oracle/_ref/deepipr_reference.zip, packed by __graft_entry__.build()
from /root/reference) unpacks and imports; the reference's own model files build from this package's blocks after
deepipr_b200.patch_reference(); its train_v23.py runs end to end (stock, CPU) through the harness; and the oracle's
restatement of the trainer step is pinned to the reference's real TrainerPrivate.train.

GPU part: INTEGRATION.md route A for real — train_v23.py / train_v1.py, unchanged, with the patched blocks on the GPU
against the same scripts stock on the CPU, and the checkpoint the patched run wrote loaded back into the stock
reference.
"""
import contextlib
import io
import json
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from oracle import ref_bundle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not ref_bundle.available(), reason="reference neither checked out nor bundled")


def _ref():
    return ref_bundle.locate(prefer_bundle=True)     # the route the GPU box takes, also where a checkout exists


def _run(code):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    p = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, env=env,
                       cwd=ROOT, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def _harness(tmp_path, tag, flavour, device, script, script_args, extra=()):
    out = tmp_path / f"{tag}.json"
    cmd = [sys.executable, "-m", "oracle.run_reference", "--script", script, "--flavour", flavour, "--device", device,
           "--out", str(out), *extra, "--", *script_args]
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900,
                       env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert p.returncode == 0, (p.stderr[-3000:], p.stdout[-500:])
    return json.load(open(out))


def test_bundle_holds_the_reference_tree():
    ref = _ref()
    for rel in ("train_v1.py", "train_v23.py", "models/layers/passportconv2d.py", "experiments/trainer_private.py",
                "passport_configs/resnet18_passport.json", "lr_configs/imagenet.json"):
        assert os.path.exists(os.path.join(ref, rel)), rel
    if os.path.isdir(os.path.join(ref_bundle.SRC, "models")):      # build container: archive == checkout, byte for byte
        for rel in ("models/layers/passportconv2d_private.py", "experiments/trainer.py", "dataset.py"):
            assert open(os.path.join(ref, rel), "rb").read() == open(os.path.join(ref_bundle.SRC, rel), "rb").read()


def test_reference_models_build_from_patched_blocks():
    REF = _ref()
    out = _run(f"""
        import sys, json, contextlib, io
        sys.path.insert(0, {ROOT!r})
        import deepipr_b200
        deepipr_b200.patch_reference()
        sys.path.insert(0, {REF!r})
        from deepipr_b200 import layers, nets
        from models.resnet_passport_private import ResNet18Private
        from models.alexnet_passport import AlexNetPassport
        import experiments.trainer_private as tp, experiments.trainer as t
        cfg = json.load(open({REF!r} + '/passport_configs/resnet18_passport.json'))
        pk = nets.passport_kwargs_from_config(cfg, 'bn', 'random', 0.1)
        with contextlib.redirect_stdout(io.StringIO()):
            m = ResNet18Private(num_classes=10, passport_kwargs=pk)
            a = AlexNetPassport(3, 10, nets.passport_kwargs_from_config(
                json.load(open({REF!r} + '/passport_configs/alexnet_passport.json')), 'bn', 'random', 0.1))
        assert type(m.layer4[0].convbnrelu_1) is layers.PassportPrivateBlock
        assert type(m.layer1[0].convbnrelu_1) is layers.ConvBlock
        assert type(a.features[4]) is layers.PassportBlock
        assert tp.PassportPrivateBlock is layers.PassportPrivateBlock and t.SignLoss is layers.SignLoss
        # same state_dict surface as the unpatched reference build
        print(json.dumps(sorted(m.state_dict().keys())))
    """)
    patched_keys = json.loads(out.strip().splitlines()[-1])
    out2 = _run(f"""
        import sys, json, contextlib, io
        sys.path.insert(0, {REF!r})
        sys.path.insert(0, {ROOT!r})
        from deepipr_b200 import nets
        from models.resnet_passport_private import ResNet18Private
        cfg = json.load(open({REF!r} + '/passport_configs/resnet18_passport.json'))
        pk = nets.passport_kwargs_from_config(cfg, 'bn', 'random', 0.1)
        with contextlib.redirect_stdout(io.StringIO()):
            m = ResNet18Private(num_classes=10, passport_kwargs=pk)
        print(json.dumps(sorted(m.state_dict().keys())))
    """)
    assert patched_keys == json.loads(out2.strip().splitlines()[-1])


def test_conv_block_patch_is_optional():
    REF = _ref()
    _run(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import deepipr_b200
        deepipr_b200.patch_reference(conv_block=False)
        sys.path.insert(0, {REF!r})
        import torch
        from models.alexnet_normal import AlexNetNormal
        m = AlexNetNormal(3, 10)
        y = m(torch.randn(2, 3, 32, 32))        # BASELINE config 1 plumbing: the reference's own CPU path still runs
        assert y.shape == (2, 10)
    """)


V3_ARGS = ["--arch", "resnet", "--batch-size", "8", "--epochs", "1", "--key-type", "shuffle", "--train-backdoor",
           "--passport-config", "passport_configs/resnet18_passport.json"]
V1_ARGS = ["--arch", "alexnet", "--batch-size", "8", "--epochs", "1", "--key-type", "random", "--train-passport",
           "--passport-config", "passport_configs/alexnet_passport.json"]


def test_reference_train_v23_runs_unmodified_on_cpu_through_the_harness(tmp_path):
    """BASELINE config 4's script (V3: private passports + trigger set), stock, CPU: the harness changes nothing but
    the data source.  Also: the patched flavour on a CPU-only machine fails loudly instead of falling back."""
    res = _harness(tmp_path, "stock", "stock", "cpu", "train_v23.py", V3_ARGS, ["--pretrained"])
    (row,) = res["history"]
    assert {"train_loss", "train_sign_loss", "train_sign_acc", "train_acc_public", "train_acc_private",
            "valid_total_acc", "wm_total_acc", "valid_s_private_layer4.1.convbn_2"} <= set(row)
    assert len(res["state_keys"]) == 147
    if not torch.cuda.is_available():
        out = tmp_path / "p.json"
        p = subprocess.run([sys.executable, "-m", "oracle.run_reference", "--script", "train_v23.py", "--flavour",
                            "patched", "--device", "cpu", "--out", str(out), "--pretrained", "--", *V3_ARGS],
                           capture_output=True, text=True, cwd=ROOT, timeout=600)
        assert p.returncode != 0 and "no CPU fallback" in p.stderr


def test_oracle_trainer_step_is_pinned_to_the_reference_trainer():
    """po.train_step / po.mirror / OracleBasicUnit against the reference's real TrainerPrivate.train and
    TesterPrivate.test_signature on the reference's real ResNet18Private (CPU, fp32), two minibatches + trigger set."""
    from deepipr_b200 import nets
    from oracle import passport_oracle as po
    from tests.helpers import quiet, seed_all
    mods = ref_bundle.import_reference(patched=False, path=_ref())
    cfg = json.load(open(os.path.join(_ref(), "passport_configs", "resnet18_passport.json")))
    pk = nets.passport_kwargs_from_config(cfg, "bn", "random", 0.1)
    seed_all(0)
    ref_model = quiet(mods["models.resnet_passport_private"].ResNet18Private, num_classes=10, passport_kwargs=pk)
    seed_all(0)
    product = quiet(nets.ResNet18, "private", 10, pk)
    g = torch.Generator().manual_seed(5)
    data = [(torch.randn(6, 3, 32, 32, generator=g), torch.randint(0, 10, (6,), generator=g)) for _ in range(2)]
    wm = [(torch.randn(2, 3, 32, 32, generator=g), torch.randint(0, 10, (2,), generator=g))]
    # lazily created random keys: create them on the reference with a fixed numpy stream, copy everything over
    import numpy as np
    np.random.seed(3)
    with torch.no_grad():
        ref_model.train()
        ref_model(data[0][0], ind=1)
    ref_model.load_state_dict({k: v.clone() for k, v in ref_model.state_dict().items()})
    product.load_state_dict(ref_model.state_dict())
    assert sorted(product.state_dict()) == sorted(ref_model.state_dict())
    oracle = po.mirror(product, round_bf16=False).train()
    opt_r = torch.optim.SGD(ref_model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    opt_o = torch.optim.SGD(oracle.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    trainer = mods["experiments.trainer_private"].TrainerPrivate(ref_model, opt_r, None, torch.device("cpu"))
    with contextlib.redirect_stdout(io.StringIO()):
        res = trainer.train(0, data, wm)
    it, ref = iter(wm), []
    for x, t in data:
        try:
            wx, wt = next(it)
        except StopIteration:
            it = iter(wm)
            wx, wt = next(it)
        ref.append(po.train_step(oracle, opt_o, torch.cat([x, wx]), torch.cat([t, wt]), private=True))
    assert abs(res["loss"] - sum(r["loss"] for r in ref) / 2) < 1e-4 * abs(res["loss"])
    assert abs(res["sign_loss"] - sum(r["sign_loss"] for r in ref)) < 1e-4 * abs(res["sign_loss"])
    assert abs(res["acc_public"] - sum(r["acc_public"] for r in ref) / 2) < 1e-4
    assert abs(res["acc_private"] - sum(r["acc_private"] for r in ref) / 2) < 1e-4
    with contextlib.redirect_stdout(io.StringIO()):
        sig = trainer.tester.test_signature()
    assert po.test_signature(oracle.eval()) == sig
    for k, v in ref_model.state_dict().items():
        if v.dtype.is_floating_point and "key" not in k:
            o = oracle.state_dict().get(k, oracle.state_dict().get(k.replace(".weight", ".conv.weight")))
            if o is not None and o.shape == v.shape:
                assert torch.allclose(o, v, rtol=1e-4, atol=2e-5), k      # two SGD steps of fp32 summation-order noise


# ------------------------------------------------------------------------------------------------ GPU: route A
def _close(a, b, rel, what):
    assert abs(a - b) <= rel * max(1.0, abs(b)), (what, a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("script,args,extra", [("train_v23.py", V3_ARGS, ["--pretrained"]),
                                               ("train_v1.py", V1_ARGS, [])], ids=["v3_resnet18", "v1_alexnet"])
def test_reference_training_scripts_run_unchanged_on_the_patched_blocks(tmp_path, script, args, extra):
    """`python train_v23.py --arch resnet --train-backdoor ...` and `python train_v1.py --train-passport ...`, unchanged:
    Experiment -> construct_model -> passport_generator.set_key -> TrainerPrivate.train / .test -> save_model.  Patched
    blocks on the GPU vs the stock reference on the CPU, same seeds and synthetic data: metrics agree to the bf16
    operand tolerance, signatures to the bit, and the checkpoint the patched run saved loads into the stock reference."""
    ckpt = tmp_path / "patched_last.pth"
    got = _harness(tmp_path, "patched", "patched", "cuda", script, args, extra + ["--save-model", str(ckpt)])
    want = _harness(tmp_path, "stock", "stock", "cpu", script, args, extra)
    assert got["library_launches"] > 100, "the patched run did not go through libpassport_sm100"
    assert got["state_keys"] == want["state_keys"]
    (g,), (w,) = got["history"], want["history"]
    assert set(g) == set(w)
    for k in w:
        if k.endswith("time") or "_time" in k:
            continue
        if "_s_" in k or k == "train_sign_acc":           # signature detection rates: bits, not tolerances
            assert abs(g[k] - w[k]) <= 2.0 / 256 + 1e-9, (k, g[k], w[k])
        elif "acc" in k:                                  # at most one image of the (small) batch decides differently
            assert abs(g[k] - w[k]) <= 100.0 / 8 + 1e-6, (k, g[k], w[k])
        elif "sign_loss" in k:
            _close(g[k], w[k], 2e-3, k)
        elif k.startswith("wm_loss"):                     # eval-mode loss on the TWO trigger images after two SGD
            _close(g[k], w[k], 0.25, k)                   # steps (running statistics barely started): chaotic
        else:
            _close(g[k], w[k], 5e-2, k)                   # cross-entropy after bf16-operand convs through 20 layers
    for k, v in want["param_abs_sums"].items():
        if "num_batches" not in k:
            _close(got["param_abs_sums"][k], v, 2e-3 if "running" not in k else 2e-2, k)
    # product -> reference checkpoint direction: the state_dict written by the patched run, loaded by the STOCK classes
    mods = ref_bundle.import_reference(patched=False, path=_ref())
    from deepipr_b200 import nets
    from tests.helpers import quiet
    if script == "train_v23.py":
        cfg = json.load(open(os.path.join(_ref(), "passport_configs", "resnet18_passport.json")))
        pk = nets.passport_kwargs_from_config(cfg, "bn", "shuffle", 0.1)
        stock = quiet(mods["models.resnet_passport_private"].ResNet18Private, num_classes=10, passport_kwargs=pk)
        tester = mods["experiments.trainer_private"].TesterPrivate(stock, torch.device("cpu"), verbose=False)
    else:
        cfg = json.load(open(os.path.join(_ref(), "passport_configs", "alexnet_passport.json")))
        pk = nets.passport_kwargs_from_config(cfg, "bn", "random", 0.1)
        stock = quiet(mods["models.alexnet_passport"].AlexNetPassport, 3, 10, pk)
        tester = None
    missing = stock.load_state_dict(torch.load(ckpt, map_location="cpu"), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    if tester is not None:
        with contextlib.redirect_stdout(io.StringIO()):
            sig = tester.test_signature()
        for k, v in sig.items():
            assert v == g["valid_s_" + k], (k, v, g["valid_s_" + k])      # fp32 reference bits == the GPU path's bits

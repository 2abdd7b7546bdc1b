"""INTEGRATION.md route A, exercised where the reference checkout exists (the build container): after
deepipr_b200.patch_reference() the reference's OWN model / trainer files import this package's blocks.
Skipped on machines without /root/reference (the GPU box)."""
import contextlib
import io
import json
import os
import subprocess
import sys
import textwrap

import pytest

REF = os.environ.get("DEEPIPR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference checkout absent")


def _run(code):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    p = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, env=env,
                       cwd=ROOT, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_models_build_from_patched_blocks():
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

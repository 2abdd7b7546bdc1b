"""Network wiring used by the benchmark, smoke test and parity tests when the reference checkout is not
available (the GPU box has no /root/reference).  Pure module plumbing — no arithmetic of its own — with the
same attribute names and therefore the same state_dict keys as the reference's

  models/resnet_passport_private.py (ResNetPrivate, :89-182)   models/resnet_passport.py (:88-180)
  models/resnet_normal.py (:52-119)                            models/alexnet_passport*.py, alexnet_normal.py

so checkpoints interchange.  One parametrised class per family instead of the reference's one file per scheme.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as F_
from .layers import ConvBlock, MaxPool2d, PassportBlock, PassportPrivateBlock

SCHEMES = ('normal', 'v1', 'private')


def passport_kwargs_from_config(config, norm_type='bn', key_type='random', sign_loss=0.1):
    """{'layer4': {'0': {'convbnrelu_1': true|false|"signature"}}, '4': true ...} -> per-block kwargs
    (what experiments/utils.py:6-50 builds from passport_configs/*.json)."""
    def leaf(flag):
        kw = {'flag': bool(flag), 'norm_type': norm_type, 'key_type': key_type, 'sign_loss': sign_loss}
        if isinstance(flag, str):
            kw['b'] = flag
        return kw

    def walk(node):
        return {k: walk(v) for k, v in node.items()} if isinstance(node, dict) else leaf(node)

    return walk(config)


def resnet18_passport_config(passport_layers=('layer4',), signature=None):
    """Same content as passport_configs/resnet18_passport.json: every conv of the listed stages is a passport layer."""
    cfg = {'convbnrelu_1': False}
    for li, name in enumerate(('layer1', 'layer2', 'layer3', 'layer4')):
        flag = name in passport_layers
        stage = {}
        for bi in range(2):
            blk = {'convbnrelu_1': flag, 'convbn_2': flag}
            if bi == 0 and li > 0:
                blk['shortcut'] = flag
            stage[str(bi)] = blk
        cfg[name] = stage
    if signature is not None and 'layer4' in passport_layers:
        cfg['layer4']['1']['convbn_2'] = signature
    return cfg


def alexnet_passport_config(passport_layers=('4', '5', '6')):
    """Same content as passport_configs/alexnet_passport.json."""
    return {k: (k in passport_layers) for k in ('0', '2', '4', '5', '6')}


class BlockSet:
    """The three block classes a net is assembled from (the tests' CPU oracle plugs its own set in here)."""

    def __init__(self, conv=ConvBlock, v1=PassportBlock, private=PassportPrivateBlock):
        self.conv, self.v1, self.private = conv, v1, private


_DEFAULT_BLOCKS = BlockSet()


def _make_block(blocks, scheme, kw, i, o, ks, s, pd, norm_type, relu=True):
    if scheme != 'normal' and kw['flag']:
        if scheme == 'private':
            return blocks.private(i, o, ks, s, pd, passport_kwargs=kw)
        return blocks.v1(i, o, ks, s, pd, passport_kwargs=kw, relu=relu)
    return blocks.conv(i, o, ks, s, pd, bn=kw['norm_type'] if kw else norm_type, relu=relu)


def _is_passport(block):
    return getattr(block, 'KIND', None) in ('v1', 'private')


def _call(block, x, force_passport, ind):
    kind = getattr(block, 'KIND', None)
    if kind == 'private':
        return block(x, force_passport, ind)
    if kind == 'v1':
        return block(x, force_passport)
    return block(x)


def _fp32_head(features, linear):
    """Classifier head in fp32 on the (bf16) activations, also under autocast: F.adaptive_avg_pool2d + nn.Linear of
    the reference nets (resnet_passport_private.py:176-179) are 0.1 % of the step, and keeping them out of bf16 means
    the logits carry no rounding beyond that of the activations they are computed from."""
    with torch.autocast('cuda', enabled=False):
        return F.linear(features.float(), linear.weight.float(), None if linear.bias is None else linear.bias.float())


class BasicUnit(nn.Module):
    """Two 3x3 blocks + (projected) shortcut; attribute names as BasicPrivateBlock / BasicPassportBlock / BasicBlock."""
    expansion = 1

    def __init__(self, scheme, in_planes, planes, stride, kwargs, norm_type, blocks=_DEFAULT_BLOCKS):
        super().__init__()
        kw = kwargs or {}
        # every block of the reference nets is built with relu=True, convbn_2 and the shortcut included
        # (resnet_normal.py:15-20, resnet_passport.py:26-30, resnet_passport_private.py:26-30)
        self.convbnrelu_1 = _make_block(blocks, scheme, kw.get('convbnrelu_1'), in_planes, planes, 3, stride, 1, norm_type)
        self.convbn_2 = _make_block(blocks, scheme, kw.get('convbn_2'), planes, planes, 3, 1, 1, norm_type)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != planes:
            self.shortcut = _make_block(blocks, scheme, kw.get('shortcut'), in_planes, planes, 1, stride, 0, norm_type)

    #: set by ResNet18: the unit's input is itself a post-ReLU tensor (the stem's or the previous unit's output), so
    #: with an identity shortcut both summands of the join are >= 0 — as they always are with a projection shortcut,
    #: whose block ends in a ReLU like every block of the reference nets.  The join relu(out + shortcut) is then a plain
    #: sum with an identity backward, which convbn_2 folds into its last pass (ConvBlock.forward(residual=)) or, for a
    #: passport convbn_2, one add launch computes.  Units used on their own keep the general relu(a + b) kernel.
    input_nonneg = False
    #: A/B switch (tests): False keeps the general relu(a + b) kernel even where the join is a plain sum
    fuse_join = True

    def _join_is_plain_sum(self):
        if not self.fuse_join:
            return False
        last_relu = getattr(self.convbn_2, 'relu', None) is not None
        if isinstance(self.shortcut, nn.Sequential):
            return last_relu and self.input_nonneg and len(self.shortcut) == 0
        return last_relu and getattr(self.shortcut, 'relu', None) is not None

    def forward(self, x, force_passport=False, ind=0):
        identity = isinstance(self.shortcut, nn.Sequential)
        c1, c2 = self.convbnrelu_1, self.convbn_2
        # identity unit made of two ConvBlocks whose join will be folded: the residual path's gradient travels from
        # convbn_2's backward to convbnrelu_1's (functional.ResidualLink) instead of being summed by autograd
        link = None
        if (F_.RESIDUAL_LINK and identity and x.is_cuda and x.requires_grad and torch.is_grad_enabled()
                and self._join_is_plain_sum()
                and isinstance(c1, ConvBlock) and isinstance(c2, ConvBlock) and c2.can_fuse_residual(x)
                and c1.conv.stride == (1, 1)):
            link = F_.ResidualLink()
            out = c1(x, _link=(link, 'add'))
        else:
            out = _call(c1, x, force_passport, ind)
        plain = self._join_is_plain_sum() and out.is_cuda
        # Folding the join into convbn_2 needs the shortcut first.  Passport blocks must run in the reference's order
        # (convbnrelu_1, convbn_2, shortcut: resnet_passport_private.py:67-85) because a block without keys draws them
        # from numpy's global RNG on its first forward (passportconv2d_private.py:198-207) — only plain ConvBlocks,
        # which consume nothing, are reordered.
        if (plain and isinstance(c2, ConvBlock) and (identity or isinstance(self.shortcut, ConvBlock))
                and c2.can_fuse_residual(out)):
            sc = x if identity else self.shortcut(x)
            if sc.dtype == out.dtype:
                return c2(out, residual=sc, _link=None if link is None else (link, 'stash'))
            out = c2(out)
        else:
            out = _call(c2, out, force_passport, ind)
            sc = x if identity else _call(self.shortcut, x, force_passport, ind)
        if plain:
            return out + sc            # relu(out + sc) == out + sc for out, sc >= 0; the gradient passes unchanged
        # F.relu(out + shortcut) (resnet_passport_private.py:78-85): one fused kernel
        return F_.add_relu(out, sc)

    def set_intermediate_keys(self, pre, x, y=None):
        def step(mine, theirs, a, b):
            if _is_passport(mine):
                mine.set_key(a, b)
            return theirs(a), (theirs(b) if b is not None else None)

        ox, oy = step(self.convbnrelu_1, pre.convbnrelu_1, x, y)
        ox, oy = step(self.convbn_2, pre.convbn_2, ox, oy)
        if isinstance(self.shortcut, nn.Sequential):
            sx, sy = x, y
        else:
            sx, sy = step(self.shortcut, pre.shortcut, x, y)
        ox = F.relu(ox + sx)
        oy = F.relu(oy + sy) if y is not None else None
        return ox, oy


class ResNet18(nn.Module):
    """ResNet-18 in the three reference flavours: scheme='normal' | 'v1' | 'private'."""

    def __init__(self, scheme='private', num_classes=10, passport_kwargs=None, norm_type='bn', imagenet=False,
                 blocks=_DEFAULT_BLOCKS):
        super().__init__()
        assert scheme in SCHEMES
        self.scheme = scheme
        pk = passport_kwargs if scheme != 'normal' else None
        if pk is None:
            pk = passport_kwargs_from_config(resnet18_passport_config(()), norm_type=norm_type)
        stem_kw = pk['convbnrelu_1']
        if num_classes == 1000 or imagenet:
            self.convbnrelu_1 = nn.Sequential(_make_block(blocks, scheme, stem_kw, 3, 64, 7, 2, 3, norm_type),
                                              MaxPool2d(3, 2, 1))
        else:
            self.convbnrelu_1 = _make_block(blocks, scheme, stem_kw, 3, 64, 3, 1, 1, norm_type)
        in_planes = 64
        for li, (planes, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)), start=1):
            units = []
            for bi, st in enumerate((stride, 1)):
                units.append(BasicUnit(scheme, in_planes, planes, st, pk[f'layer{li}'][str(bi)], norm_type, blocks))
                units[-1].input_nonneg = True      # fed by the stem (ReLU) or by the previous unit's join
                in_planes = planes
            setattr(self, f'layer{li}', nn.Sequential(*units))
        self.linear = nn.Linear(512, num_classes)
        #: Opt-in common-subexpression elimination for the V2/V3 step (trainer_private.py:159-161 calls the model
        #: twice on the SAME batch, ind=0 then ind=1): every block in front of the first passport layer computes
        #: identical values in both calls, so the second call can reuse the first call's trunk output (and autograd
        #: then back-propagates the summed gradient through the trunk once).  Results are those of the two full
        #: passes up to summation order; BatchNorm running statistics receive their second EMA update explicitly.
        #: Off by default: bench.py reports it separately (value_shared_trunk).
        self.share_trunk = False
        self._trunk_cache = None

    def _stages(self):
        return (self.layer1, self.layer2, self.layer3, self.layer4)

    def _units(self):
        return [u for stage in self._stages() for u in stage]

    def _trunk_len(self):
        """Number of leading basic units (after a passport-free stem) that contain no passport block."""
        stem = self.convbnrelu_1[0] if isinstance(self.convbnrelu_1, nn.Sequential) else self.convbnrelu_1
        if _is_passport(stem):
            return -1
        n = 0
        for u in self._units():
            if any(_is_passport(m) for m in u.modules()):
                break
            n += 1
        return n

    def _stem(self, x, force_passport, ind):
        # blocks keep the memory format of their input: enter the NHWC pipeline once, on the 3-channel image
        x = x.contiguous(memory_format=torch.channels_last)
        if isinstance(self.convbnrelu_1, nn.Sequential):
            return self.convbnrelu_1[1](_call(self.convbnrelu_1[0], x, force_passport, ind))
        return _call(self.convbnrelu_1, x, force_passport, ind)

    def _trunk_bns(self, ntrunk):
        mods = [self.convbnrelu_1] + self._units()[:ntrunk]
        return [m for root in mods for m in root.modules() if isinstance(m, nn.BatchNorm2d)]

    def _trunk(self, x, ntrunk):
        """Stem + the first `ntrunk` units, memoised on the identity of the input batch and of the weights."""
        params = [p for root in [self.convbnrelu_1] + self._units()[:ntrunk] for p in root.parameters()]
        key = (id(x), x.data_ptr(), x._version, tuple(x.shape), x.dtype, torch.is_grad_enabled(), self.training,
               torch.is_autocast_enabled(), F_.weight_epoch(), sum(p._version for p in params))
        cache = self._trunk_cache
        if cache is not None and cache[0] == key:
            _, out, bns, before = cache
            if self.training and bns:
                # second EMA update with the same batch statistics s:  r1 = (1-m) r0 + m s  =>
                # r2 = (1-m) r1 + m s = (2-m) r1 - (1-m) r0      (momentum m per BatchNorm, default 0.1)
                with torch.no_grad():
                    for bn, (m0, v0) in zip(bns, before):
                        mom = bn.momentum
                        # .data: stock BatchNorm (the CPU oracle mirror of this wiring) keeps the running buffers in
                        # its autograd graph and would reject a version bump before backward
                        bn.running_mean.data.mul_(2.0 - mom).sub_(m0, alpha=1.0 - mom)
                        bn.running_var.data.mul_(2.0 - mom).sub_(v0, alpha=1.0 - mom)
                        bn.num_batches_tracked.data.add_(1)
            self._trunk_cache = None        # one reuse per batch: public pass -> private pass
            return out
        bns = self._trunk_bns(ntrunk) if self.training else []
        before = [(bn.running_mean.clone(), bn.running_var.clone()) for bn in bns]
        out = self._stem(x, False, 0)
        for unit in self._units()[:ntrunk]:
            out = unit(out, False, 0)
        self._trunk_cache = (key, out, bns, before)
        return out

    def forward(self, x, force_passport=False, ind=0):
        ntrunk = self._trunk_len() if self.share_trunk else -1
        if ntrunk >= 0:
            out = self._trunk(x, ntrunk)
            rest = self._units()[ntrunk:]
        else:
            out = self._stem(x, force_passport, ind)
            rest = self._units()
        for unit in rest:
            out = unit(out, force_passport, ind)
        return _fp32_head(out.float().mean(dim=(2, 3)), self.linear)

    def set_intermediate_keys(self, pretrained_model, x, y=None):
        """Push passport candidates through a (normal) pretrained net, handing each passport layer its input
        activations as keys (reference resnet_passport_private.py:146-162)."""
        with torch.no_grad():
            stem = self.convbnrelu_1[0] if isinstance(self.convbnrelu_1, nn.Sequential) else self.convbnrelu_1
            if _is_passport(stem):
                stem.set_key(x, y)
            x = pretrained_model.convbnrelu_1(x)
            y = pretrained_model.convbnrelu_1(y) if y is not None else None
            for mine, theirs in zip(self._stages(), pretrained_model._stages()):
                for unit, pre in zip(mine, theirs):
                    x, y = unit.set_intermediate_keys(pre, x, y)


_ALEX_OUT = {0: 64, 2: 192, 4: 384, 5: 256, 6: 256}
_ALEX_KP = {0: (5, 2), 2: (5, 2), 4: (3, 1), 5: (3, 1), 6: (3, 1)}


class AlexNetCifar(nn.Module):
    """CIFAR AlexNet (features idx 0..7, max-pool at 1,3,7, Linear(4096->classes)); scheme as above."""

    def __init__(self, scheme='v1', in_channels=3, num_classes=10, passport_kwargs=None, norm_type='bn',
                 blocks=_DEFAULT_BLOCKS):
        super().__init__()
        assert scheme in SCHEMES
        self.scheme = scheme
        if passport_kwargs is None or scheme == 'normal':
            passport_kwargs = passport_kwargs_from_config(alexnet_passport_config(()), norm_type=norm_type)
        layers, inp = [], in_channels
        for idx in range(8):
            if idx in (1, 3, 7):
                layers.append(MaxPool2d(2, 2))
                continue
            k, p = _ALEX_KP[idx]
            layers.append(_make_block(blocks, scheme, passport_kwargs[str(idx)], inp, _ALEX_OUT[idx], k, 1, p, norm_type))
            inp = _ALEX_OUT[idx]
        self.features = nn.Sequential(*layers)
        self.classifier = nn.Linear(4 * 4 * 256, num_classes)

    def forward(self, x, force_passport=False, ind=0):
        x = x.contiguous(memory_format=torch.channels_last)
        for m in self.features:
            x = _call(m, x, force_passport, ind)
        return _fp32_head(x.float().reshape(x.size(0), -1), self.classifier)      # logical NCHW flatten order

    def set_intermediate_keys(self, pretrained_model, x, y=None):
        with torch.no_grad():
            for theirs, mine in zip(pretrained_model.features, self.features):
                if _is_passport(mine):
                    mine.set_key(x, y)
                x = theirs(x)
                y = theirs(y) if y is not None else None

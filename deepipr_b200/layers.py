"""Host-side mirror of the reference's passport layer surface, backed by libpassport_sm100.

Same class names, constructor signatures, attributes, buffers, state_dict keys and RNG consumption order as

  models/layers/passportconv2d.py          PassportBlock           (reference :11-223)
  models/layers/passportconv2d_private.py  PassportPrivateBlock    (reference :11-219)
  models/layers/conv2d.py                  ConvBlock               (reference :5-36)
  models/losses/sign_loss.py               SignLoss                (reference :6-63)

so the reference's models/, experiments/ and attack scripts run unchanged on top of them
(see deepipr_b200.patch_reference and INTEGRATION.md).  The arithmetic is NOT PyTorch's: forward and
backward go through deepipr_b200.functional (C ABI -> sm_100a kernels).  CPU tensors raise.
"""
import random

import numpy as np
import torch
import torch.nn as nn
import torch.nn.init as init

from . import _lib as L
from . import functional as F_


class SignLoss(nn.Module):
    """Hinge sign loss on the passport scale (reference sign_loss.py:6-63).

    Stateful protocol kept verbatim: ``reset()`` zeroes ``loss``/``acc`` (python ints), ``add(scale)``
    accumulates tensors into them, trainers read ``m.loss`` / ``m.acc``.
    """

    def __init__(self, alpha, b=None):
        super().__init__()
        self.alpha = alpha
        self.register_buffer('b', b)
        self.loss = 0
        self.acc = 0
        self.scale_cache = None

    def set_b(self, b):
        self.b.copy_(b)

    def _need_cache(self):
        if self.scale_cache is None:
            raise Exception('scale_cache is None')
        return self.scale_cache

    def get_acc(self):
        # mean(sign(b) == sign(scale))  (sign_loss.py:18-23) — bookkeeping helper, off the hot path
        scale = self._need_cache()
        return (torch.sign(self.b.view(-1)) == torch.sign(scale.view(-1))).float().mean()

    def get_loss(self):
        # alpha * sum(relu(0.1 - b*scale))  (sign_loss.py:25-30) — the hinge term only
        scale = self._need_cache()
        return (self.alpha * torch.relu(-self.b.view(-1) * scale.view(-1) + 0.1)).sum()

    def add(self, scale):
        """loss += hinge + 1e-5*sum(scale^2); acc += sign agreement (sign_loss.py:32-54), one fused kernel."""
        loss, acc = F_.sign_loss(scale, self.b, self.alpha)
        self._add_fused(scale, loss, acc)

    def _add_fused(self, scale, loss, acc):
        self.scale_cache = scale
        self.loss += loss
        self.acc += acc

    def reset(self):
        self.loss = 0
        self.acc = 0
        self.scale_cache = None


def _norm_mode(bn_module):
    """Which fused norm variant a block's ``bn`` attribute maps to (None => run it as a torch module)."""
    if isinstance(bn_module, nn.BatchNorm2d):
        if bn_module.training or not bn_module.track_running_stats:
            return L.PP_NORM_BN_TRAIN
        return L.PP_NORM_BN_EVAL
    if bn_module is None or (isinstance(bn_module, nn.Sequential) and len(bn_module) == 0):
        return L.PP_NORM_NONE
    if isinstance(bn_module, nn.GroupNorm):
        return L.PP_NORM_GN
    if isinstance(bn_module, nn.InstanceNorm2d) and not bn_module.track_running_stats:
        return L.PP_NORM_GN          # instance norm == group norm with one channel per group
    return None


def _norm_groups(bn_module):
    if isinstance(bn_module, nn.GroupNorm):
        return bn_module.num_groups
    if isinstance(bn_module, nn.InstanceNorm2d):
        return bn_module.num_features
    return 0


#: package-wide default arithmetic of the contractions; a block's own ``precision`` attribute overrides it
_DEFAULT_PRECISION = 'bf16'


def set_precision(precision):
    """Arithmetic of every block that does not set its own ``precision``:

    ``'bf16'``  bf16 activations, tcgen05 kind::f16 — BASELINE configs 3-5 (the model runs under autocast(bf16) or is
                fed bf16); fp32 inputs are converted on entry and the output returns in the input's dtype.
    ``'tf32'``  fp32 activations end to end, tcgen05 kind::tf32 — BASELINE config 2 (train_v1.py:13-29 runs the
                reference in fp32, where torch's cuDNN convolutions use TF32 by default).  Applies to fp32 inputs
                outside autocast; bf16 inputs, autocast regions and group / instance norm blocks keep the bf16 path.
    Returns the previous setting."""
    global _DEFAULT_PRECISION
    if precision not in ('bf16', 'tf32'):
        raise ValueError(f"precision must be 'bf16' or 'tf32', got {precision!r}")
    prev, _DEFAULT_PRECISION = _DEFAULT_PRECISION, precision
    return prev


def get_precision():
    return _DEFAULT_PRECISION


class _FusedConvMixin:
    """Weight-operand cache + dispatch shared by the three block types."""

    #: 'bf16' | 'tf32' | None (package default, see set_precision)
    precision = None

    def _dtype(self, x, norm):
        """PP_DTYPE_* of this call (see set_precision)."""
        p = self.precision or _DEFAULT_PRECISION
        if (p == 'tf32' and x.dtype == torch.float32 and not torch.is_autocast_enabled()
                and norm != L.PP_NORM_GN):
            return L.PP_DTYPE_TF32
        return L.PP_DTYPE_BF16

    #: keep the conv output z (saved for backward, re-read by the affine pass) in fp32.  bf16 halves that
    #: traffic but adds a second rounding in front of the bf16 output (DESIGN.md "Precision").
    z_f32 = True

    def _spec(self):
        conv = self.conv
        return F_.ConvSpec(C=conv.in_channels, O=conv.out_channels, kh=conv.kernel_size[0], kw=conv.kernel_size[1],
                           stride=conv.stride[0], pad=conv.padding[0])

    def _check_conv(self):
        conv = self.conv
        if conv.stride[0] != conv.stride[1] or conv.padding[0] != conv.padding[1] or conv.dilation != (1, 1) \
                or conv.groups != 1 or conv.padding_mode != 'zeros':
            raise RuntimeError("deepipr_b200: only square stride/padding, dilation 1, groups 1 are supported")

    def invalidate_cache(self):
        """Drop cached bf16 weight operands / pooled keys (call after mutating ``weight.data`` or keys in place)."""
        self.__dict__.pop('_pp_prepared', None)
        self.__dict__.pop('_pp_keypool', None)

    def _prepared(self, dtype=L.PP_DTYPE_BF16):
        """Operand copies of the conv weight (pp_weight_prep): bf16, or fp32 re-layouts for PP_DTYPE_TF32.

        Tensor._version does not see edits made through ``.data`` (the idiom of the reference's pruning / flip attack
        scripts: ``p.data.mul_(mask)``, ``w.data.copy_(...)``), so the cached copies are trusted only where every
        writer is known: the parameter lives in a parallel.FlatParams (FlatSGD bumps the weight epoch on each step),
        the module is in training mode and autograd is recording.  Everywhere else — evaluation, attack scripts,
        stock optimizers — the copies are rebuilt on every forward (one tiny kernel)."""
        w = self.conv.weight
        cached = self.__dict__.get('_pp_prepared')
        if (cached is not None and torch.is_grad_enabled() and self.training
                and getattr(w, '_pp_flat_slot', None) is not None
                and cached.version == w._version and cached.data_ptr == w.data_ptr()
                and cached.epoch == F_.weight_epoch() and cached.wf.device == w.device and cached.dtype == dtype):
            return cached
        prepared = F_.prepare_weight(w, self._spec(), need_dgrad=True, dtype=dtype)
        self.__dict__['_pp_prepared'] = prepared
        return prepared

    def _bn_opts(self, norm, relu, z_f32, x, dtype=L.PP_DTYPE_BF16):
        bn = getattr(self, 'bn', None)
        rm = rv = None
        eps, momentum = 1e-5, 0.1
        if isinstance(bn, nn.BatchNorm2d):
            eps = bn.eps
            if bn.momentum is None:
                raise RuntimeError("deepipr_b200: BatchNorm momentum=None (cumulative average) is not supported")
            momentum = bn.momentum
            rm, rv = bn.running_mean, bn.running_var
            if norm == L.PP_NORM_BN_TRAIN and bn.training and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
            if norm == L.PP_NORM_BN_TRAIN and not bn.training:
                rm = rv = None  # track_running_stats=False in eval: batch statistics, nothing to update
        groups = 0
        if norm == L.PP_NORM_GN:
            eps, groups = bn.eps, _norm_groups(bn)
            if isinstance(bn, nn.InstanceNorm2d):
                spec = self._spec()
                P, Q = spec.out_hw(x.shape[2], x.shape[3])
                if P * Q == 1 and bn.training:   # same refusal as F.instance_norm (torch/nn/functional.py)
                    raise ValueError(f"Expected more than 1 spatial element when training, got input size "
                                     f"{torch.Size((x.shape[0], spec.O, P, Q))}")
        out_dtype = torch.bfloat16 if (torch.is_autocast_enabled() or x.dtype == torch.bfloat16) else x.dtype
        return F_.BlockOpts(spec=self._spec(), norm=norm, relu=bool(relu),
                            z_f32=bool(z_f32) or dtype == L.PP_DTYPE_TF32, eps=float(eps),
                            momentum=float(momentum), running_mean=rm, running_var=rv, out_dtype=out_dtype,
                            groups=int(groups), dtype=dtype)


class ConvBlock(nn.Module, _FusedConvMixin):
    """conv -> BN/GN/IN (affine) -> ReLU (reference conv2d.py:5-36)."""
    KIND = 'conv'

    def __init__(self, i, o, ks=3, s=1, pd=1, bn='bn', relu=True):
        super().__init__()
        self.conv = nn.Conv2d(i, o, ks, s, pd, bias=bn == 'none')
        if bn == 'bn':
            self.bn = nn.BatchNorm2d(o)
        elif bn == 'gn':
            self.bn = nn.GroupNorm(o // 16, o)
        elif bn == 'in':
            self.bn = nn.InstanceNorm2d(o)
        else:
            self.bn = None
        self.relu = nn.ReLU(inplace=True) if relu else None
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_normal_(self.conv.weight, mode='fan_out', nonlinearity='relu')

    def can_fuse_residual(self, x):
        """True when forward(x, residual=...) can fold the residual join of a basic unit into this block's last pass
        (bf16 arithmetic, batch-norm or no norm, a call that keeps z: training-mode BN or autograd recording)."""
        norm = _norm_mode(self.bn)
        if norm not in (L.PP_NORM_NONE, L.PP_NORM_BN_TRAIN, L.PP_NORM_BN_EVAL) or self.relu is None:
            return False
        if self._dtype(x, norm) != L.PP_DTYPE_BF16:
            return False
        return norm == L.PP_NORM_BN_TRAIN or torch.is_grad_enabled()

    def forward(self, x, residual=None, _link=None):
        """``residual`` (extension over the reference signature, used by nets.BasicUnit): a tensor >= 0 of the output's
        shape that is added after the ReLU, i.e. the unit's residual join folded into this block
        (F.relu(out + shortcut) of resnet_passport_private.py:78-85 with both summands already non-negative).
        ``_link``: (functional.ResidualLink, role) — the two ConvBlocks of a unit hand the residual path's gradient
        from one backward to the other instead of leaving the sum to autograd."""
        F_.require_cuda(x, "ConvBlock input")
        self._check_conv()
        norm = _norm_mode(self.bn)
        dtype = self._dtype(x, norm)
        prepared = self._prepared(dtype)
        if norm is None:
            # a norm module this library has no kernel for (e.g. InstanceNorm with running statistics):
            # fused conv, then the module itself
            o = self._bn_opts(L.PP_NORM_NONE, False, False, x, dtype)
            y = F_.conv_block(x, self.conv.weight, None, self.conv.bias, prepared, o)
            y = self.bn(y)
            return self.relu(y) if self.relu is not None else y
        if norm == L.PP_NORM_NONE:
            gamma, beta = None, self.conv.bias
        else:
            # BatchNorm2d / GroupNorm carry an affine; InstanceNorm2d(o) does not (weight, bias are None)
            gamma, beta = self.bn.weight, self.bn.bias
        o = self._bn_opts(norm, self.relu is not None, self.z_f32, x, dtype)
        o.direct_grad_ok = True      # conv.weight / bn.weight / bn.bias (or conv.bias) feed this operator only
        o.link = _link
        if residual is not None:
            if not self.can_fuse_residual(x):
                raise RuntimeError("deepipr_b200: this ConvBlock call cannot fuse a residual (see can_fuse_residual)")
        return F_.conv_block(x, self.conv.weight, gamma, beta, prepared, o, residual)


class MaxPool2d(nn.MaxPool2d):
    """nn.MaxPool2d whose channels_last CUDA inputs take the library's NHWC kernels (pp_maxpool_fwd / _bwd: one pass
    over the bytes each way, argmax kept as one byte per output); anything else — CPU tensors, NCHW memory, dilation,
    ceil_mode, return_indices — is the stock module.  Same constructor, no parameters, same state_dict."""

    def forward(self, x):
        def one(v):
            return v if isinstance(v, int) else (v[0] if v[0] == v[1] else None)

        k, s, p, d = one(self.kernel_size), one(self.stride), one(self.padding), one(self.dilation)
        if (None not in (k, s, p) and d == 1 and not self.ceil_mode and not self.return_indices
                and F_.max_pool2d_supported(x, k, s, p)):
            return F_.max_pool2d(x, k, s, p)
        return super().forward(x)


def _signature_bits(b, o):
    """passport_kwargs['b'] -> +-1 tensor of length o (reference passportconv2d.py:25-40)."""
    if isinstance(b, int):
        return torch.ones(o) * b
    if isinstance(b, str):
        if len(b) * 8 > o:
            raise Exception('Too much bit information')
        bits = torch.sign(torch.rand(o) - 0.5)
        pos = 0
        for ch in b:
            for bit in format(ord(ch), 'b').zfill(8):
                bits[pos] = -1 if bit == '0' else 1
                pos += 1
        return bits
    return b


class _PassportBase(nn.Module, _FusedConvMixin):
    """Everything V1 (PassportBlock) and V2/V3 (PassportPrivateBlock) share.

    Subclasses define the buffer names of the passport (``key``/``skey`` vs ``key_private``/``skey_private``)
    and which SignLoss attribute the passport path feeds.
    """

    _KEY = 'key'
    _SKEY = 'skey'

    def _build(self, i, o, ks, s, pd, passport_kwargs):
        if passport_kwargs == {}:
            print('Warning, passport_kwargs is empty')
        self.conv = nn.Conv2d(i, o, ks, s, pd, bias=False)
        self.key_type = passport_kwargs.get('key_type', 'random')
        self.weight = self.conv.weight
        self.alpha = passport_kwargs.get('sign_loss', 1)
        # the default is evaluated eagerly, exactly like dict.get(..., default) in the reference (RNG order!)
        b = passport_kwargs.get('b', torch.sign(torch.rand(o) - 0.5))
        self.register_buffer('b', _signature_bits(b, o))
        self.requires_reset_key = False

    def _build_norm(self, o, norm_type):
        if norm_type == 'bn':
            self.bn = nn.BatchNorm2d(o, affine=False)
        elif norm_type == 'gn':
            self.bn = nn.GroupNorm(o // 16, o, affine=False)
        elif norm_type == 'in':
            self.bn = nn.InstanceNorm2d(o, affine=False)
        else:
            self.bn = nn.Sequential()

    # ---- learnable public affine (reference passportconv2d.py:73-87)
    def init_bias(self, force_init=False):
        if force_init:
            self.bias = nn.Parameter(torch.Tensor(self.conv.out_channels).to(self.weight.device))
            init.zeros_(self.bias)
        else:
            self.bias = None

    def init_scale(self, force_init=False):
        if force_init:
            self.scale = nn.Parameter(torch.Tensor(self.conv.out_channels).to(self.weight.device))
            init.ones_(self.scale)
        else:
            self.scale = None

    def reset_parameters(self):
        init.kaiming_normal_(self.weight, mode='fan_out', nonlinearity='relu')

    # ---- passport construction (reference passportconv2d.py:90-137, 198-207); host-side, one-time
    def passport_selection(self, passport_candidates):
        b, c, h, w = passport_candidates.size()
        if c == 3:  # network input: take one whole image
            return passport_candidates[random.randint(0, b - 1)].unsqueeze(0)
        flat = passport_candidates.contiguous().view(b * c, h, w)     # (also accepts channels_last candidates)
        taken = [False] * (b * c)
        chosen = []
        img = 0
        while len(chosen) < c:
            if img >= b:
                img = 0
            pick = img * c + random.randint(0, c - 1)
            while taken[pick]:
                pick = img * c + random.randint(0, c - 1)
            taken[pick] = True
            chosen.append(flat[pick].unsqueeze(0).unsqueeze(0))
            img += 1
        return torch.cat(chosen, dim=1)

    def set_key(self, x, y=None):
        if int(x.size(0)) != 1:
            x = self.passport_selection(x)
            if y is not None:
                y = self.passport_selection(y)
        self.register_buffer(self._KEY, x)
        self.register_buffer(self._SKEY, y)
        self.__dict__.pop('_pp_keypool', None)

    def generate_key(self, *shape):
        newshape = list(shape)
        newshape[0] = 1
        return np.random.uniform(-1.0, 1.0, newshape)

    def get_scale_key(self):
        return getattr(self, self._SKEY)

    def get_bias_key(self):
        return getattr(self, self._KEY)

    def _maybe_random_key(self, x):
        if (getattr(self, self._KEY) is None and self.key_type == 'random') or self.requires_reset_key:
            self.set_key(torch.tensor(self.generate_key(*x.size()), dtype=x.dtype, device=x.device),
                         torch.tensor(self.generate_key(*x.size()), dtype=x.dtype, device=x.device))

    # ---- passport-derived affine
    def _pooled_keys(self):
        key, skey = getattr(self, self._KEY), getattr(self, self._SKEY)
        if key is None or skey is None:
            raise RuntimeError("deepipr_b200: passport key/skey not set (call set_key or use key_type='random')")
        F_.require_cuda(key, "passport key")
        sig = (key.data_ptr(), key._version, tuple(key.shape), skey.data_ptr(), skey._version, tuple(skey.shape),
               str(key.device))
        cached = self.__dict__.get('_pp_keypool')
        # as for the weights: ``key.data.copy_()`` is invisible to _version, so outside a recording training step the
        # pooled keys are rebuilt every time (two tiny kernels)
        if cached is not None and cached[0] == sig and torch.is_grad_enabled() and self.training:
            return cached[1], cached[2]
        spec = self._spec()
        S_skey, S_key = F_.key_pool(skey, spec), F_.key_pool(key, spec)
        self.__dict__['_pp_keypool'] = (sig, S_skey, S_key)
        return S_skey, S_key

    def _passport_affine(self, loss_module):
        """(gamma, beta) from the passport; feeds ``loss_module`` exactly as get_scale does in the reference."""
        self._check_conv()
        key, skey = getattr(self, self._KEY), getattr(self, self._SKEY)
        S_skey, S_key = self._pooled_keys()
        b = loss_module.b if loss_module is not None else None
        alpha = loss_module.alpha if loss_module is not None else 0.0
        # passport_attack_3.py turns the keys into Parameters: differentiate through the pooled-key identity
        keys_need_grad = torch.is_grad_enabled() and (key.requires_grad or skey.requires_grad)
        actx = F_.AffineCtx(self._spec(), S_skey, S_key,
                            None if b is None else b.detach().reshape(-1).float().contiguous(), float(alpha),
                            tuple(key.shape) if keys_need_grad else None)
        if keys_need_grad:
            if key.shape != skey.shape:
                raise RuntimeError("deepipr_b200: key and skey must have the same shape to be optimised")
            gamma, beta, loss, acc = F_.passport_affine(self.weight, actx, skey, key)
        else:
            gamma, beta, loss, acc = F_.passport_affine(self.weight, actx)
        if loss_module is not None:
            loss_module.reset()
            loss_module._add_fused(gamma.view(1, -1, 1, 1), loss, acc)
        return gamma, beta

    def _run_passport(self, x, loss_module, relu):
        """The whole passport block as one operator (F_.passport_conv -> pp_passport_conv_fwd): gamma / beta from the
        passport, the sign loss fed to ``loss_module`` exactly as get_scale() does, conv, norm, affine, ReLU.
        Returns None when this call needs the composed path instead (keys being optimised, a norm module without a
        kernel here)."""
        F_.require_cuda(x, "passport block input")
        self._check_conv()
        norm = _norm_mode(self.bn)
        key, skey = getattr(self, self._KEY), getattr(self, self._SKEY)
        if norm is None or key is None or skey is None:
            return None
        if torch.is_grad_enabled() and (key.requires_grad or skey.requires_grad):
            return None
        S_skey, S_key = self._pooled_keys()
        b = loss_module.b if loss_module is not None else None
        pc = F_.PassportCtx(S_skey, S_key, None if b is None else b.detach().reshape(-1).float().contiguous(),
                            float(loss_module.alpha) if loss_module is not None else 0.0)
        dtype = self._dtype(x, norm)
        o = self._bn_opts(norm, relu, self.z_f32, x, dtype)
        y, gamma, beta, loss, acc = F_.passport_conv(x, self.weight, self._prepared(dtype), o, pc)
        if loss_module is not None:
            loss_module.reset()
            loss_module._add_fused(gamma.view(1, -1, 1, 1), loss, acc)
        return y

    def _run(self, x, gamma, beta, relu):
        """conv -> norm -> gamma*x+beta -> relu with per-channel gamma/beta tensors of O elements."""
        F_.require_cuda(x, "passport block input")
        self._check_conv()
        norm = _norm_mode(self.bn)
        dtype = self._dtype(x, norm)
        prepared = self._prepared(dtype)
        if norm is None:   # norm module without a kernel here: fused conv, then the module and the affine in torch
            o = self._bn_opts(L.PP_NORM_NONE, False, False, x, dtype)
            y = F_.conv_block(x, self.weight, None, None, prepared, o)
            y = self.bn(y)
            y = gamma.view(1, -1, 1, 1).to(y.dtype) * y + beta.view(1, -1, 1, 1).to(y.dtype)
            return torch.relu_(y) if relu else y
        o = self._bn_opts(norm, relu, self.z_f32, x, dtype)
        return F_.conv_block(x, self.weight, gamma, beta, prepared, o)

    def _load_placeholders(self, state_dict, prefix):
        """Pre-allocate key / scale / bias slots so the default loader can copy into them
        (reference passportconv2d.py:177-196)."""
        for name in (self._KEY, self._SKEY):
            if prefix + name in state_dict:
                self.register_buffer(name, torch.randn(*state_dict[prefix + name].size()))
        if prefix + 'scale' in state_dict:
            self.scale = nn.Parameter(torch.randn(*state_dict[prefix + 'scale'].size()))
        if prefix + 'bias' in state_dict:
            self.bias = nn.Parameter(torch.randn(*state_dict[prefix + 'bias'].size()))
        self.invalidate_cache()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        self._load_placeholders(state_dict, prefix)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)


class PassportBlock(_PassportBase):
    """V1 passport block (reference passportconv2d.py:11-223)."""
    KIND = 'v1'

    def __init__(self, i, o, ks=3, s=1, pd=1, passport_kwargs={}, relu=True):
        super().__init__()
        self._build(i, o, ks, s, pd, passport_kwargs)
        self.sign_loss = SignLoss(self.alpha, self.b) if self.alpha != 0 else None
        self.register_buffer('key', None)
        self.register_buffer('skey', None)
        self.init_scale()
        self.init_bias()
        self._build_norm(o, passport_kwargs.get('norm_type', 'bn'))
        self.relu = nn.ReLU(inplace=True) if relu else None
        self.reset_parameters()

    def _affine(self, force_passport):
        """(gamma, beta) as [O] tensors; passport path is evaluated once for both."""
        use_scale = self.scale is not None and not force_passport
        use_bias = self.bias is not None and not force_passport
        if use_scale and use_bias:
            return self.scale, self.bias
        gamma, beta = self._passport_affine(self.sign_loss if not use_scale else None)
        return (self.scale if use_scale else gamma), (self.bias if use_bias else beta)

    def get_scale(self, force_passport=False):
        if self.scale is not None and not force_passport:
            return self.scale.view(1, -1, 1, 1)
        gamma, _ = self._passport_affine(self.sign_loss)
        return gamma.view(1, -1, 1, 1)

    def get_bias(self, force_passport=False):
        if self.bias is not None and not force_passport:
            return self.bias.view(1, -1, 1, 1)
        _, beta = self._passport_affine(None)
        return beta.view(1, -1, 1, 1)

    def forward(self, x, force_passport=False):
        self._maybe_random_key(x)
        use_scale = self.scale is not None and not force_passport
        use_bias = self.bias is not None and not force_passport
        if not use_scale and not use_bias:                 # the plain passport path: one fused operator
            y = self._run_passport(x, self.sign_loss, self.relu is not None)
            if y is not None:
                return y
        gamma, beta = self._affine(force_passport)
        return self._run(x, gamma, beta, self.relu is not None)


class PassportPrivateBlock(_PassportBase):
    """V2/V3 block: public learnable scale/bias (ind=0) and private passport-derived ones (ind=1)
    (reference passportconv2d_private.py:11-219)."""

    KIND = 'private'
    _KEY = 'key_private'
    _SKEY = 'skey_private'

    def __init__(self, i, o, ks=3, s=1, pd=1, passport_kwargs={}):
        super().__init__()
        self._build(i, o, ks, s, pd, passport_kwargs)
        self.norm_type = passport_kwargs.get('norm_type', 'bn')
        self.init_public_bit = passport_kwargs.get('init_public_bit', True)
        self.sign_loss_private = SignLoss(self.alpha, self.b)
        self.register_buffer('key_private', None)
        self.register_buffer('skey_private', None)
        self.init_scale(True)
        self.init_bias(True)
        self._build_norm(o, self.norm_type)
        self.relu = nn.ReLU(inplace=True)
        self.reset_parameters()

    def _affine(self, force_passport, ind):
        use_scale = self.scale is not None and not force_passport and ind == 0
        use_bias = self.bias is not None and not force_passport and ind == 0
        if use_scale and use_bias:
            return self.scale, self.bias
        gamma, beta = self._passport_affine(self.sign_loss_private if not use_scale else None)
        return (self.scale if use_scale else gamma), (self.bias if use_bias else beta)

    def get_scale(self, force_passport=False, ind=0):
        if self.scale is not None and not force_passport and ind == 0:
            return self.scale.view(1, -1, 1, 1)
        gamma, _ = self._passport_affine(self.sign_loss_private)
        return gamma.view(1, -1, 1, 1)

    def get_bias(self, force_passport=False, ind=0):
        if self.bias is not None and not force_passport and ind == 0:
            return self.bias.view(1, -1, 1, 1)
        _, beta = self._passport_affine(None)
        return beta.view(1, -1, 1, 1)

    def forward(self, x, force_passport=False, ind=0):
        self._maybe_random_key(x)
        use_scale = self.scale is not None and not force_passport and ind == 0
        use_bias = self.bias is not None and not force_passport and ind == 0
        if not use_scale and not use_bias:                 # the private passport path: one fused operator
            y = self._run_passport(x, self.sign_loss_private, True)
            if y is not None:
                return y
        gamma, beta = self._affine(force_passport, ind)
        return self._run(x, gamma, beta, True)

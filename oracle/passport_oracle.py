"""CPU oracle for the passport-layer hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file;
the product package (deepipr_b200/) never does.

What it is: a restatement, in plain PyTorch CPU ops, of the algorithm of kamwoh/DeepIPR's passport blocks.  The
arithmetic itself lives in a third-party dependency of the reference that is not vendored in its tree: PyTorch
(requirements.txt:1 `torch`, unpinned; Dockerfile:1 pins pytorch 1.8.0; this container has torch 2.11.0).  The
oracle therefore calls the same published operators the reference's call sites call —

    F.conv2d        <- nn.Conv2d          models/layers/passportconv2d.py:18,148,169,218
    F.batch_norm    <- nn.BatchNorm2d     models/layers/passportconv2d.py:58,219   (affine=False)
    Tensor.mean     <- GAP of the key     models/layers/passportconv2d.py:151-152,172-173
    F.relu          <- ReLU / hinge       models/layers/passportconv2d.py:222, models/losses/sign_loss.py:27

— and restates everything around them (block composition, gamma/beta derivation, SignLoss accumulation,
public/private switch, trainer step) from the reference lines cited on each function.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against outputs
of the reference itself, generated in the build container by tests/golden/make_golden.py (imports
/root/reference read-only) and committed under tests/golden/*.pt; tests/test_oracle_golden.py checks every
function here against them.

`round_bf16=True` makes every block round the operands of its batch convolution (input activation, weight) to
bf16 before the fp32 computation — the "fp32 accumulate on bf16-rounded operands" model the CUDA path is
compared with (parity definition).  The passport affine (gamma, beta = GAP(conv(W, skey / key)))
is NOT rounded: the CUDA path evaluates it from the fp32 master weight and the fp32 keys, because sign(gamma) is
the signature and has to match the reference's fp32 get_scale() bit for bit.  Block outputs and residual joins ARE
rounded in this mode — activations are bf16 tensors on the CUDA path ("bf16 activations", BASELINE.json north_star),
each produced by one rounding of an fp32 result — so that whole-network logits can be held to 1e-3.

`round_bf16='tf32'` is the model of BASELINE config 2 (AlexNet V1 in fp32, train_v1.py:13-29): the reference's fp32
modules on a GPU run their cuDNN convolutions in TF32 (torch.backends.cudnn.allow_tf32 defaults to True), i.e. every
operand of the three contractions of a convolution — (x, W) forward, (dz, W) data gradient, (x, dz) weight gradient —
is cut to 10 explicit mantissa bits and accumulated in fp32; everything else (normalisation, affine, losses, the
passport affine) stays fp32.  Nothing is rounded to bf16 in this mode.
"""
import copy

import torch
import torch.nn as nn
import torch.nn.functional as F


def bf16_round(t, enabled=True):
    """Round to bf16 and back, straight-through for autograd."""
    if not enabled:
        return t
    r = t.detach().to(torch.bfloat16).to(t.dtype)
    return t + (r - t.detach()) if t.requires_grad else r


def tf32_cut(t):
    """The value a TF32 tensor-core instruction reads from an fp32 operand: the low 13 mantissa bits are dropped."""
    return (t.detach().contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


class _ConvTF32(torch.autograd.Function):
    """nn.Conv2d forward / backward with TF32 operands and fp32 accumulation (what cuDNN runs for the reference's fp32
    nn.Conv2d on Ampere-and-later GPUs; passportconv2d.py:18,218)."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, pad):
        ctx.save_for_backward(x, w)
        ctx.sp = (stride, pad)
        ctx.has_bias = bias is not None
        return F.conv2d(tf32_cut(x), tf32_cut(w), None if bias is None else bias.detach(), stride, pad)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        stride, pad = ctx.sp
        gc = tf32_cut(g)
        dx = torch.nn.grad.conv2d_input(x.shape, tf32_cut(w), gc, stride, pad) if ctx.needs_input_grad[0] else None
        dw = torch.nn.grad.conv2d_weight(tf32_cut(x), w.shape, gc, stride, pad) if ctx.needs_input_grad[1] else None
        db = g.sum((0, 2, 3)) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db, None, None


def batch_conv(x, weight, bias, stride, pad, mode=False):
    """The block's convolution over the minibatch under an operand model: False = plain fp32, True = operands rounded
    to bf16, 'tf32' = TF32 operands in all three contractions."""
    if mode == 'tf32':
        return _ConvTF32.apply(x, weight, bias, stride, pad)
    rb = mode is True
    return F.conv2d(bf16_round(x, rb), bf16_round(weight, rb), bias, stride, pad)


# ------------------------------------------------------------------------------------------------
# functional restatement
# ------------------------------------------------------------------------------------------------
def key_affine(weight, key, stride, pad):
    """GAP(conv(W, key)) -> [O]   (passportconv2d.py:146-152 for skey->scale, :167-173 for key->bias)."""
    out = F.conv2d(key, weight, None, stride, pad)           # [Bk, O, P, Q]
    bk, o = out.size(0), out.size(1)
    return out.view(bk, o, -1).mean(dim=2).mean(dim=0)       # mean over positions, then over the key batch


def sign_loss_terms(scale, b, alpha):
    """SignLoss.add: returns (loss increment, acc increment)  (sign_loss.py:18-30, 53-54)."""
    s = scale.reshape(-1)
    bb = b.reshape(-1)
    hinge = (alpha * F.relu(-bb * s + 0.1)).sum()
    reg = 0.00001 * s.pow(2).sum()
    acc = (torch.sign(bb) == torch.sign(s)).float().mean()
    return hinge + reg, acc


def normalise(z, kind, running_mean=None, running_var=None, training=True, momentum=0.1, eps=1e-5, groups=None):
    """The block's `bn` attribute, affine=False (passportconv2d.py:56-64)."""
    if kind == 'bn':
        return F.batch_norm(z, running_mean, running_var, None, None, training, momentum, eps)
    if kind == 'gn':
        return F.group_norm(z, groups, None, None, eps)
    if kind == 'in':
        return F.instance_norm(z, eps=eps)
    return z


def passport_forward(x, weight, gamma, beta, stride, pad, norm_kind, relu, mode=False, **norm_args):
    """y = relu(gamma * norm(conv(x, W)) + beta)   (passportconv2d.py:218-222)."""
    z = batch_conv(x, weight, None, stride, pad, mode)
    zn = normalise(z, norm_kind, **norm_args)
    y = gamma.view(1, -1, 1, 1) * zn + beta.view(1, -1, 1, 1)
    return F.relu(y) if relu else y


# ------------------------------------------------------------------------------------------------
# module-shaped oracle: mirrors a product net block by block so whole networks can be compared / timed
# ------------------------------------------------------------------------------------------------
class OracleSignLoss(nn.Module):
    """sign_loss.py:6-63."""

    def __init__(self, alpha, b):
        super().__init__()
        self.alpha = alpha
        self.register_buffer('b', b)
        self.reset()

    def reset(self):
        self.loss, self.acc, self.scale_cache = 0, 0, None

    def add(self, scale):
        self.scale_cache = scale
        loss, acc = sign_loss_terms(scale, self.b, self.alpha)
        self.loss += loss
        self.acc += acc


class _OracleBlockBase(nn.Module):
    round_bf16 = False

    def _conv_args(self):
        return self.conv.stride, self.conv.padding

    def _norm_kind(self):
        bn = getattr(self, 'bn', None)
        if isinstance(bn, nn.BatchNorm2d):
            return 'bn'
        if isinstance(bn, nn.GroupNorm):
            return 'gn'
        if isinstance(bn, nn.InstanceNorm2d):
            return 'in'
        return 'none'

    def _norm_args(self):
        bn = getattr(self, 'bn', None)
        kind = self._norm_kind()
        if kind == 'bn':
            if bn.training and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
            return dict(running_mean=bn.running_mean, running_var=bn.running_var, training=bn.training,
                        momentum=bn.momentum, eps=bn.eps)
        if kind == 'gn':
            return dict(groups=bn.num_groups, eps=bn.eps)
        if kind == 'in':
            return dict(eps=bn.eps)
        return {}


class OracleConvBlock(_OracleBlockBase):
    """conv2d.py:5-36 — conv -> norm(affine) -> relu."""
    KIND = 'conv'

    @classmethod
    def from_product(cls, m, round_bf16):
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.conv = copy.deepcopy(m.conv)
        self.bn = copy.deepcopy(m.bn)
        self.has_relu = m.relu is not None
        self.round_bf16 = round_bf16
        return self

    def forward(self, x):
        stride, pad = self._conv_args()
        z = batch_conv(x, self.conv.weight, self.conv.bias, stride, pad, self.round_bf16)
        kind = self._norm_kind()
        if kind != 'none':
            zn = normalise(z, kind, **self._norm_args())
            if self.bn.weight is not None:
                zn = zn * self.bn.weight.view(1, -1, 1, 1) + self.bn.bias.view(1, -1, 1, 1)
            z = zn
        return bf16_round(F.relu(z) if self.has_relu else z, self.round_bf16 is True)


class OraclePassportBlock(_OracleBlockBase):
    """passportconv2d.py:11-223 (V1) and passportconv2d_private.py:11-219 (private=True)."""

    @classmethod
    def from_product(cls, m, round_bf16):
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.private = m.KIND == 'private'
        self.KIND = m.KIND
        self.conv = copy.deepcopy(m.conv)
        self.weight = self.conv.weight
        self.bn = copy.deepcopy(m.bn)
        self.has_relu = m.relu is not None
        self.alpha = m.alpha
        self.register_buffer('b', m.b.detach().clone())
        key, skey = m.get_bias_key(), m.get_scale_key()
        self.register_buffer('key', None if key is None else key.detach().clone())
        self.register_buffer('skey', None if skey is None else skey.detach().clone())
        self.scale = None if m.scale is None else nn.Parameter(m.scale.detach().clone())
        self.bias = None if m.bias is None else nn.Parameter(m.bias.detach().clone())
        src_loss = m.sign_loss_private if self.private else m.sign_loss
        self.sign_loss = None if src_loss is None else OracleSignLoss(src_loss.alpha, self.b)
        self.round_bf16 = round_bf16
        return self

    def get_scale(self, force_passport=False, ind=0):
        use_public = self.scale is not None and not force_passport and (ind == 0 or not self.private)
        if use_public:
            return self.scale.view(1, -1, 1, 1)
        stride, pad = self._conv_args()
        scale = key_affine(self.weight, self.skey, stride, pad).view(1, -1, 1, 1)     # never rounded (see header)
        if self.sign_loss is not None:
            self.sign_loss.reset()
            self.sign_loss.add(scale)
        return scale

    def get_bias(self, force_passport=False, ind=0):
        use_public = self.bias is not None and not force_passport and (ind == 0 or not self.private)
        if use_public:
            return self.bias.view(1, -1, 1, 1)
        stride, pad = self._conv_args()
        return key_affine(self.weight, self.key, stride, pad).view(1, -1, 1, 1)

    def forward(self, x, force_passport=False, ind=0):
        stride, pad = self._conv_args()
        gamma = self.get_scale(force_passport, ind)
        beta = self.get_bias(force_passport, ind)
        y = passport_forward(x, self.weight, gamma.reshape(-1), beta.reshape(-1), stride, pad, self._norm_kind(),
                             self.has_relu, mode=self.round_bf16, **self._norm_args())
        return bf16_round(y, self.round_bf16 is True)


def _call_block(block, x, force_passport, ind):
    kind = getattr(block, 'KIND', None)
    if kind == 'private':
        return block(x, force_passport, ind)
    if kind == 'v1':
        return block(x, force_passport)
    return block(x)


class OracleBasicUnit(nn.Module):
    """Residual unit: BasicPrivateBlock.forward (models/resnet_passport_private.py:67-86), BasicPassportBlock.forward
    (models/resnet_passport.py:66-85), BasicBlock.forward (models/resnet_normal.py:22-27):
    out = relu(convbn_2(convbnrelu_1(x)) + shortcut(x))."""

    @classmethod
    def from_children(cls, convbnrelu_1, convbn_2, shortcut, round_bf16):
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.convbnrelu_1, self.convbn_2, self.shortcut = convbnrelu_1, convbn_2, shortcut
        self.round_bf16 = round_bf16
        return self

    def forward(self, x, force_passport=False, ind=0):
        out = _call_block(self.convbnrelu_1, x, force_passport, ind)
        out = _call_block(self.convbn_2, out, force_passport, ind)
        if isinstance(self.shortcut, nn.Sequential):
            sc = x
            for m in self.shortcut:          # empty for the identity shortcut
                sc = m(sc)
        else:
            sc = _call_block(self.shortcut, x, force_passport, ind)
        return bf16_round(F.relu(out + sc), self.round_bf16 is True)


def mirror(model, round_bf16=False):
    """Deep-copy a product network to the CPU, swapping every fused block for its oracle restatement."""
    def convert(mod):
        kind = getattr(mod, 'KIND', None)
        if kind == 'conv':
            return OracleConvBlock.from_product(mod, round_bf16)
        if kind in ('v1', 'private'):
            return OraclePassportBlock.from_product(mod, round_bf16)
        return None

    def cpu_copy(mod):
        # drop the product's device-side caches before copying
        for m in mod.modules():
            if hasattr(m, 'invalidate_cache'):
                m.invalidate_cache()
        return copy.deepcopy(mod).cpu()

    root = convert(model)
    if root is not None:
        return root.cpu()
    out = cpu_copy(model)

    def walk(parent):
        for name, child in list(parent.named_children()):
            repl = convert(child)
            if repl is not None:
                setattr(parent, name, repl)
            else:
                walk(child)
                kids = dict(child.named_children())
                if {'convbnrelu_1', 'convbn_2', 'shortcut'} <= set(kids):      # a residual unit of the product wiring
                    setattr(parent, name, OracleBasicUnit.from_children(kids['convbnrelu_1'], kids['convbn_2'],
                                                                        kids['shortcut'], round_bf16))

    walk(out)
    return out.cpu()


def sign_loss_modules(model):
    return [m for m in model.modules() if isinstance(m, OracleSignLoss)]


def accuracy_top1(output, target):
    """experiments/trainer.py:28-43 with topk=(1,)."""
    with torch.no_grad():
        pred = output.argmax(dim=1)
        return pred.eq(target).float().sum().mul_(100.0 / target.size(0))


def train_step(model, optimizer, data, target, private):
    """One minibatch of Trainer.train (experiments/trainer.py:128-148, private=False) or TrainerPrivate.train
    (experiments/trainer_private.py:148-177, private=True).  Works on product nets and on mirrors alike.
    Returns dict(loss, sign_loss, acc...)."""
    from_types = (OracleSignLoss,)
    try:  # product SignLoss is only importable when the package is; the oracle must also work stand-alone
        from deepipr_b200.layers import SignLoss as _ProductSignLoss
        from_types = (OracleSignLoss, _ProductSignLoss)
    except Exception:
        pass
    optimizer.zero_grad()
    losses = [m for m in model.modules() if isinstance(m, from_types)]
    for m in losses:
        m.reset()
    out = {}
    if private:
        loss = torch.zeros((), device=data.device)
        for ind in range(2):
            pred = model(data, ind=ind)
            loss = loss + F.cross_entropy(pred, target)
            out['acc_public' if ind == 0 else 'acc_private'] = accuracy_top1(pred, target).item()
    else:
        pred = model(data)
        loss = F.cross_entropy(pred, target)
    sign_loss = torch.zeros((), device=data.device)
    for m in losses:
        sign_loss = sign_loss + m.loss
    (loss + sign_loss).backward()
    optimizer.step()
    out['sign_loss'] = float(sign_loss)
    out['loss'] = float(loss)
    if not private:
        out['acc'] = accuracy_top1(pred, target).item()
    return out


def test_signature(model):
    """TesterPrivate.test_signature (experiments/trainer_private.py:37-71): per-layer fraction of matching bits."""
    res = {}
    with torch.no_grad():
        for name, m in model.named_modules():
            kind = getattr(m, 'KIND', None)
            if kind == 'private':
                bits = m.get_scale(ind=1).view(-1).sign()
                res['private_' + name] = (bits == m.b).float().mean().item()
            elif kind == 'v1':
                bits = m.get_scale().view(-1).sign()
                res['public_' + name] = (bits == m.b).float().mean().item()
    return res

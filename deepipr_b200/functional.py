"""Autograd bindings of the C ABI: the two differentiable operators the passport blocks are made of.

  conv_block        y = relu?(gamma * norm(conv(x, W)) + beta)           (pp_conv_block_fwd / _bwd)
  passport_affine   (gamma, beta, sign_loss, sign_acc) from W and the pooled passport keys
                                                                         (pp_passport_affine_fwd / _bwd)

Tensors cross the boundary as raw device pointers on torch's current stream.  Everything here requires
CUDA tensors; CPU tensors raise (the package has no CPU path by design).
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib as L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"deepipr_b200: {what} is on {t.device}; the passport kernels are sm_100a CUDA only (no CPU fallback)")
    # the library launches on the calling thread's CURRENT device and on torch's current stream of that device
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(
            f"deepipr_b200: {what} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
            "run one process per GPU (torchrun) or wrap the call in torch.cuda.device(...) — nn.DataParallel-style "
            "multi-device threads are not supported")


@dataclass(frozen=True)
class ConvSpec:
    """Static geometry of a block (constructor arguments of the reference blocks)."""
    C: int
    O: int
    kh: int
    kw: int
    stride: int
    pad: int

    def out_hw(self, H, W):
        return ((H + 2 * self.pad - self.kh) // self.stride + 1, (W + 2 * self.pad - self.kw) // self.stride + 1)


_desc_cache = {}
_ws_bytes_cache = {}
_workspace = {}
#: PP_ALGO_* used by every call; tests flip it to compare the tensor-core path with the SIMT path
ALGO = L.PP_ALGO_AUTO
#: bumped whenever parameters are rewritten behind autograd's back (FlatSGD / FlatParams), so that cached bf16
#: weight operands keyed on Tensor._version are refreshed
_weight_epoch = 0


def bump_weight_epoch():
    global _weight_epoch
    _weight_epoch += 1


def weight_epoch():
    return _weight_epoch


def make_desc(spec: ConvSpec, N, H, W, norm=L.PP_NORM_NONE, relu=0, z_f32=0, eps=1e-5, momentum=0.1, algo=None,
              groups=0, flags=0, dtype=L.PP_DTYPE_BF16):
    algo = ALGO if algo is None else algo
    key = (spec, N, H, W, norm, relu, z_f32, eps, momentum, algo, groups, flags, dtype)
    d = _desc_cache.get(key)
    if d is None:
        d = L.PPConvDesc(N=N, C=spec.C, H=H, W=W, O=spec.O, kh=spec.kh, kw=spec.kw, stride=spec.stride, pad=spec.pad,
                         norm=norm, relu=int(relu), z_f32=int(z_f32), eps=eps, momentum=momentum, algo=algo,
                         groups=int(groups), flags=int(flags), dtype=int(dtype))
        _desc_cache[key] = d
        if len(_desc_cache) > 4096:
            _desc_cache.clear()
    return d, key


def workspace(desc, key, which, device, slot=0):
    """Grow-only per-device scratch buffer (fwd and bwd share it: they are stream-ordered).  slot 1 is the scratch of
    the weight-gradient launches that run on the side stream (see _SideStream)."""
    ck = (key, which)
    nbytes = _ws_bytes_cache.get(ck)
    if nbytes is None:
        out = C.c_size_t(0)
        L.check(L.load().pp_workspace_bytes(C.byref(desc), which, C.byref(out)), "pp_workspace_bytes")
        nbytes = int(out.value)
        _ws_bytes_cache[ck] = nbytes
    wk = device if slot == 0 else (device, slot)
    buf = _workspace.get(wk)
    if buf is None or buf.numel() < nbytes:
        if buf is not None and slot != 0:
            join_side(device)             # the old side buffer may still be in use by a launch in flight
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspace[wk] = buf
    return buf, nbytes


# ---------------------------------------------------------------- weight gradients on a second stream
#: The weight gradient of a block is a tensor-core kernel that nothing later in the backward pass depends on, while the
#: next block's backward starts with two HBM-bound passes (reduce, dz).  Where the gradient is accumulated by the
#: kernels themselves into a flat gradient buffer (FlatParams, direct accumulation: autograd never sees it), it is
#: launched on a side stream so that the two overlap (measured: 50-70 % of the shorter kernel is hidden).  Everything
#: that reads the gradients joins first: the end of the backward pass (autograd engine callback), a bucket all-reduce,
#: the optimizer step.  PP_NO_WGRAD_OVERLAP=1 (or functional.OVERLAP_WGRAD = False) keeps everything on one stream.
import os as _os
OVERLAP_WGRAD = _os.environ.get("PP_NO_WGRAD_OVERLAP", "0") != "1"


class _SideStream:
    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.pending = False        # launches since the last join
        self.keep = []              # tensors the side launches read / write, kept alive until the join
        self.cb_queued = False


_side = {}


def _side_for(device):
    st = _side.get(device)
    if st is None:
        st = _side[device] = _SideStream(device)
    return st


def join_side(device=None):
    """Make the current stream wait for the weight-gradient launches in flight on the side stream of `device` (all
    devices if None) and release the tensors they used."""
    for dev, st in list(_side.items()):
        if device is not None and dev != device:
            continue
        st.cb_queued = False
        if st.pending:
            with torch.cuda.device(dev):
                torch.cuda.current_stream(dev).wait_stream(st.stream)
            st.pending = False
        st.keep.clear()


def to_nhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    """Logical NCHW tensor whose memory is dense NHWC bf16 (no copy if it already is)."""
    if x.dtype != torch.bfloat16:
        x = x.to(torch.bfloat16)
    return x.contiguous(memory_format=torch.channels_last)


def act_dtype(dtype: int) -> torch.dtype:
    """Element type of x / y / dy / dx for a PP_DTYPE_* arithmetic type."""
    return torch.float32 if dtype == L.PP_DTYPE_TF32 else torch.bfloat16


def to_nhwc(x: torch.Tensor, dtype: int) -> torch.Tensor:
    """Dense NHWC memory in the activation type of `dtype` (bf16, or fp32 for PP_DTYPE_TF32); no copy if it already is."""
    t = act_dtype(dtype)
    if x.dtype != t:
        x = x.to(t)
    return x.contiguous(memory_format=torch.channels_last)


@dataclass
class PreparedWeight:
    wf: torch.Tensor                 # [O, kh, kw, C], bf16 (fp32 for PP_DTYPE_TF32)
    wd: Optional[torch.Tensor]       # [C, kh, kw, O] or None
    version: int = -1
    data_ptr: int = 0
    epoch: int = -1
    dtype: int = L.PP_DTYPE_BF16


def prepare_weight(weight: torch.Tensor, spec: ConvSpec, need_dgrad: bool, dtype: int = L.PP_DTYPE_BF16) -> PreparedWeight:
    """fp32 OIHW master weight -> operand copies in the two layouts the tensor-core kernels read (pp_weight_prep)."""
    require_cuda(weight, "conv weight")
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    t = act_dtype(dtype)
    wf = torch.empty((spec.O, spec.kh, spec.kw, spec.C), dtype=t, device=w.device)
    wd = torch.empty((spec.C, spec.kh, spec.kw, spec.O), dtype=t, device=w.device) if need_dgrad else None
    d, _ = make_desc(spec, 1, max(spec.kh, 1), max(spec.kw, 1), dtype=dtype)
    L.check(L.load().pp_weight_prep(C.byref(d), L.ptr(w), L.ptr(wf), L.ptr(wd), _stream()), "pp_weight_prep")
    return PreparedWeight(wf, wd, weight._version, weight.data_ptr(), _weight_epoch, dtype)


def master_weight(weight: torch.Tensor) -> torch.Tensor:
    """The fp32 OIHW master weight as a dense tensor (no copy for an ordinary nn.Conv2d parameter)."""
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    return w


def key_pool(key: torch.Tensor, spec: ConvSpec) -> torch.Tensor:
    """Pooled passport patch S (fp64 [kh*kw*C]) such that GAP(conv(W, key)) == W.view(O, -1) @ S."""
    require_cuda(key, "passport key")
    k = key.detach()
    if k.dtype != torch.float32 or not k.is_contiguous():
        k = k.float().contiguous()
    Bk, Ck, H, W = k.shape
    if Ck != spec.C:
        raise RuntimeError(f"passport key has {Ck} channels, block expects {spec.C}")
    S = torch.empty(spec.kh * spec.kw * spec.C, dtype=torch.float64, device=k.device)
    d, _ = make_desc(spec, 1, H, W)
    L.check(L.load().pp_key_pool(C.byref(d), int(Bk), L.ptr(k), L.ptr(S), _stream()), "pp_key_pool")
    return S


@dataclass
class AffineCtx:
    spec: ConvSpec
    S_skey: torch.Tensor
    S_key: torch.Tensor
    b: Optional[torch.Tensor]
    alpha: float
    key_shape: Optional[tuple] = None      # (Bk, C, H, W) when gradients w.r.t. the keys are wanted


class _PassportAffineFn(torch.autograd.Function):
    """gamma = GAP(conv(W, skey)), beta = GAP(conv(W, key)) + SignLoss.add(gamma)
    (reference: passportconv2d.py:142-175, sign_loss.py:32-54)."""

    @staticmethod
    def forward(ctx, weight, actx: AffineCtx, skey=None, key=None):
        dev = weight.device
        O = actx.spec.O
        # the fp32 OIHW master weight itself (not the bf16 operand copy): sign(gamma) must be the reference's
        w = master_weight(weight)
        gamma = torch.empty(O, dtype=torch.float32, device=dev)
        beta = torch.empty(O, dtype=torch.float32, device=dev)
        has_b = actx.b is not None
        loss = torch.zeros((), dtype=torch.float32, device=dev) if has_b else None
        acc = torch.zeros((), dtype=torch.float32, device=dev) if has_b else None
        d, _ = make_desc(actx.spec, 1, actx.spec.kh, actx.spec.kw)
        L.check(L.load().pp_passport_affine_fwd(
            C.byref(d), L.ptr(w), L.ptr(actx.S_skey), L.ptr(actx.S_key), L.ptr(actx.b),
            float(actx.alpha), L.ptr(gamma), L.ptr(beta), L.ptr(loss), L.ptr(acc), _stream()),
            "pp_passport_affine_fwd")
        ctx.actx = actx
        ctx.save_for_backward(gamma, w if actx.key_shape is not None else None)
        ctx.wshape = weight.shape
        if has_b:
            ctx.mark_non_differentiable(acc)
            return gamma, beta, loss, acc
        return gamma, beta

    @staticmethod
    def backward(ctx, g_gamma, g_beta, g_loss=None, g_acc=None):
        actx = ctx.actx
        gamma, w = ctx.saved_tensors
        dev = gamma.device
        dw = torch.empty(ctx.wshape, dtype=torch.float32, device=dev)

        def f32(t):
            return None if t is None else t.contiguous().float()

        g_gamma, g_beta, g_loss = f32(g_gamma), f32(g_beta), f32(g_loss)
        d, _ = make_desc(actx.spec, 1, actx.spec.kh, actx.spec.kw)
        if ctx.needs_input_grad[0]:
            L.check(L.load().pp_passport_affine_bwd(
                C.byref(d), L.ptr(actx.S_skey), L.ptr(actx.S_key), L.ptr(gamma), L.ptr(actx.b), float(actx.alpha),
                L.ptr(g_gamma), L.ptr(g_beta), L.ptr(g_loss), L.ptr(dw), 0, _stream()), "pp_passport_affine_bwd")
        else:
            dw = None
        dskey = dkey = None
        want_s = len(ctx.needs_input_grad) > 2 and ctx.needs_input_grad[2]
        want_k = len(ctx.needs_input_grad) > 3 and ctx.needs_input_grad[3]
        if want_s or want_k:
            Bk, Ck, H, W = actx.key_shape
            dk, _ = make_desc(actx.spec, 1, H, W)
            scratch = torch.empty(2 * actx.spec.kh * actx.spec.kw * actx.spec.C, dtype=torch.float64, device=dev)
            dskey = torch.empty(actx.key_shape, dtype=torch.float32, device=dev) if want_s else None
            dkey = torch.empty(actx.key_shape, dtype=torch.float32, device=dev) if want_k else None
            L.check(L.load().pp_passport_key_grad(
                C.byref(dk), int(Bk), L.ptr(w), L.ptr(gamma), L.ptr(actx.b), float(actx.alpha),
                L.ptr(g_gamma), L.ptr(g_beta), L.ptr(g_loss), L.ptr(scratch), L.ptr(dskey), L.ptr(dkey), _stream()),
                "pp_passport_key_grad")
        return dw, None, dskey, dkey


def passport_affine(weight, actx: AffineCtx, skey=None, key=None):
    """Returns (gamma[O], beta[O], sign_loss or None, sign_acc or None).  Pass `skey` / `key` only when they
    require grad (they are then differentiated through the pooled-key identity)."""
    require_cuda(weight, "conv weight")
    out = _PassportAffineFn.apply(weight, actx, skey, key)
    if len(out) == 2:
        return out[0], out[1], None, None
    return out


def signature_verify(entries, want_gamma=False):
    """Batched ownership verification: ``entries`` = [(weight fp32 [O,C,kh,kw], S_skey fp64 [K], b fp32 [O])] for
    every passport layer (all on one device).  Returns (matched int32 [n] device tensor, O list, gamma list or None);
    detection of layer i = matched[i] / O[i]  (trainer_private.py:37-71).  One kernel launch per 64 layers."""
    if not entries:
        return None, [], ([] if want_gamma else None)
    dev = entries[0][1].device
    n = len(entries)
    matched = torch.empty(n, dtype=torch.int32, device=dev)
    Os = [int(b.numel()) for _, _, b in entries]
    gamma = torch.empty(sum(Os), dtype=torch.float32, device=dev) if want_gamma else None
    keep, ofs = [], 0
    for start in range(0, n, L.PP_SIG_MAX_LAYERS):
        chunk = entries[start:start + L.PP_SIG_MAX_LAYERS]
        arr = (L.PPSigLayer * len(chunk))()
        for i, (weight, S, b) in enumerate(chunk):
            require_cuda(S, "pooled skey")
            require_cuda(weight, "conv weight")
            bf = b.detach().reshape(-1).float().contiguous()
            w = master_weight(weight)
            keep.extend((bf, w))
            O, K = w.shape[0], w.numel() // w.shape[0]
            if S.numel() != K or bf.numel() != O:
                raise RuntimeError(f"signature_verify: layer {start + i} has O={O} K={K} but |S|={S.numel()} |b|={bf.numel()}")
            arr[i] = L.PPSigLayer(w.data_ptr(), S.data_ptr(), bf.data_ptr(), O, K, ofs, int(w.shape[1]))
            ofs += O
        L.check(L.load().pp_signature_verify(len(chunk), arr, C.c_void_p(matched[start:].data_ptr()), L.ptr(gamma),
                                             _stream()), "pp_signature_verify")
    gammas = list(torch.split(gamma, Os)) if want_gamma else None
    return matched, Os, gammas


class _SignLossFn(torch.autograd.Function):
    """SignLoss.add on an arbitrary scale tensor (sign_loss.py:18-54)."""

    @staticmethod
    def forward(ctx, scale, b, alpha):
        g = scale.detach().reshape(-1).float().contiguous()
        bb = b.detach().reshape(-1).float().contiguous()
        loss = torch.zeros((), dtype=torch.float32, device=g.device)
        acc = torch.zeros((), dtype=torch.float32, device=g.device)
        L.check(L.load().pp_sign_loss_fwd(int(g.numel()), L.ptr(g), L.ptr(bb), float(alpha), L.ptr(loss), L.ptr(acc),
                                          _stream()), "pp_sign_loss_fwd")
        ctx.save_for_backward(g, bb)
        ctx.alpha = float(alpha)
        ctx.shape = scale.shape
        ctx.dtype = scale.dtype
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    def backward(ctx, g_loss, g_acc=None):
        g, bb = ctx.saved_tensors
        gg = torch.empty_like(g)
        gl = g_loss.contiguous().float()
        L.check(L.load().pp_sign_loss_bwd(int(g.numel()), L.ptr(g), L.ptr(bb), ctx.alpha, L.ptr(gl), L.ptr(gg),
                                          _stream()), "pp_sign_loss_bwd")
        return gg.reshape(ctx.shape).to(ctx.dtype), None, None


def sign_loss(scale, b, alpha):
    require_cuda(scale, "scale")
    return _SignLossFn.apply(scale, b, alpha)


class _CeTop1Fn(torch.autograd.Function):
    """F.cross_entropy(pred, target) + accuracy(pred, target)[0] and the logits gradient in one launch
    (experiments/trainer_private.py:161-168, experiments/trainer.py:28-43)."""

    @staticmethod
    def forward(ctx, logits, target):
        lg = logits.detach()
        if lg.dtype not in (torch.float32, torch.bfloat16):
            lg = lg.float()
        lg = lg.contiguous()
        tg = target.detach().to(torch.int64).contiguous()
        N, classes = lg.shape
        out = torch.empty(2, dtype=torch.float32, device=lg.device)
        need = ctx.needs_input_grad[0]
        dl = torch.empty((N, classes), dtype=torch.float32, device=lg.device) if need else None
        L.check(L.load().pp_ce_top1(int(N), int(classes), L.ptr(lg), int(lg.dtype == torch.bfloat16), L.ptr(tg),
                                    C.c_void_p(out.data_ptr()), C.c_void_p(out.data_ptr() + 4), L.ptr(dl), 0,
                                    _stream()), "pp_ce_top1")
        ctx.save_for_backward(dl)
        ctx.dtype = logits.dtype
        loss, top1 = out[0], out[1]
        ctx.mark_non_differentiable(top1)
        return loss, top1

    @staticmethod
    def backward(ctx, g_loss, g_top1=None):
        (dl,) = ctx.saved_tensors
        return (dl * g_loss).to(ctx.dtype), None


def ce_top1(logits, target):
    """(mean cross-entropy, precision@1 in percent) of [N, classes] logits — one fused kernel, no host read."""
    require_cuda(logits, "logits")
    if logits.dim() != 2:
        raise RuntimeError(f"ce_top1 expects [N, classes] logits, got {tuple(logits.shape)}")
    return _CeTop1Fn.apply(logits, target)


@dataclass
class BlockOpts:
    spec: ConvSpec
    norm: int                 # PP_NORM_*
    relu: bool
    z_f32: bool
    eps: float = 1e-5
    momentum: float = 0.1
    running_mean: Optional[torch.Tensor] = None
    running_var: Optional[torch.Tensor] = None
    out_dtype: Optional[torch.dtype] = None
    algo: Optional[int] = None
    groups: int = 0           # PP_NORM_GN: number of groups (== O for InstanceNorm)
    dtype: int = L.PP_DTYPE_BF16   # PP_DTYPE_*: bf16 tensors / kind::f16, or fp32 tensors / kind::tf32
    #: (ResidualLink, 'stash' | 'add') or None — see ResidualLink
    link: Optional[tuple] = None
    #: the block's weight / gamma / beta are Parameters that receive gradients from this operator ONLY (ConvBlock):
    #: when they live in a parallel.FlatParams their gradients are accumulated straight into the flat buffer by
    #: the producing kernels (PP_FLAG_ACC_*), and autograd sees no gradient for them
    direct_grad_ok: bool = False


#: PP_RESIDUAL_LINK=1 folds the gradient sum at the input of a residual unit into conv1's data-gradient epilogue
#: (ResidualLink below).  OFF by default: measured on B200 the epilogue's loads of the second summand are exposed
#: (the 32 loads of a chunk are ordered behind the chunk's stores and a chunk's latency is not hidden by the next
#: tile's main loop at the 64-channel geometry), which costs far more than the 0.3 ms of autograd's add launches it
#: removes — 29.3 ms per step against 15.2 ms.  Kept for the parity tests that pin its numerics (bit-identical sums).
RESIDUAL_LINK = _os.environ.get("PP_RESIDUAL_LINK", "0") == "1"


class ResidualLink:
    """Carries the gradient of a unit's residual path from the block that closes the unit to the block that opens it.

    x -> conv1 -> conv2(+ x) -> y: autograd would hand `gy` to x twice — through conv1's data gradient and through the
    residual — and sum the two in a separate pass.  With a link, conv2's backward (role 'stash') parks gy here and
    reports no gradient for the residual; conv1's backward (role 'add'), which always runs later, has it added in its
    data-gradient kernel's epilogue (pp_conv_block_bwd_dz dx_add) and returns the complete gradient of x."""
    __slots__ = ("g",)

    def __init__(self):
        self.g = None


def _flat_slot(t):
    """(FlatParams, index) when `t` is a Parameter re-homed by parallel.FlatParams with direct accumulation on."""
    slot = getattr(t, '_pp_flat_slot', None)
    if slot is None:
        return None
    flat = slot[0]()
    if flat is None or not flat.direct or flat.params[slot[1]] is not t:
        return None
    # only while .grad IS the flat view (an external zero_grad(set_to_none=True) or a foreign optimizer takes the
    # parameter back to the ordinary autograd path)
    if t.grad is None or t.grad.data_ptr() != flat.grad_view(slot[1]).data_ptr():
        return None
    return flat, slot[1]


class _ConvBlockFn(torch.autograd.Function):
    """Fused conv -> (batch-norm) -> per-channel affine -> ReLU and its backward
    (reference: passportconv2d.py:218-222, passportconv2d_private.py:215-218, conv2d.py:29-36)."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, prepared: PreparedWeight, o: BlockOpts, residual=None):
        spec = o.spec
        N, Cx, H, W = x.shape
        if Cx != spec.C:
            raise RuntimeError(f"input has {Cx} channels, block expects {spec.C}")
        P, Q = spec.out_hw(H, W)
        dev = x.device
        if prepared.dtype != o.dtype:
            raise RuntimeError("deepipr_b200: weight operands were prepared for another arithmetic type")
        adt = act_dtype(o.dtype)
        xc = to_nhwc(x.detach(), o.dtype)
        need_grad = any(ctx.needs_input_grad[:4]) or ctx.needs_input_grad[6]
        keep_z = need_grad or o.norm in (L.PP_NORM_BN_TRAIN, L.PP_NORM_GN) or residual is not None
        d, key = make_desc(spec, N, H, W, o.norm, o.relu, o.z_f32, o.eps, o.momentum, o.algo, o.groups, dtype=o.dtype)
        y = torch.empty((N, spec.O, P, Q), dtype=adt, device=dev, memory_format=torch.channels_last)
        z = torch.empty((N, P, Q, spec.O), dtype=torch.float32 if o.z_f32 else torch.bfloat16, device=dev) \
            if keep_z else None
        nstat = N * o.groups if o.norm == L.PP_NORM_GN else spec.O   # GN/IN: one (mean, invstd) per sample and group
        save_mean = torch.empty(nstat, dtype=torch.float32, device=dev)
        save_invstd = torch.empty(nstat, dtype=torch.float32, device=dev)
        g = None if gamma is None else gamma.detach().reshape(-1).float().contiguous()
        b = None if beta is None else beta.detach().reshape(-1).float().contiguous()
        ws, nbytes = workspace(d, key, L.PP_WS_FWD, dev)
        res = None
        if residual is not None:
            # residual join folded into the block's last pass: y = bf16(relu(...)) + residual (see pp_conv_block_fwd_res)
            if z is None or o.dtype != L.PP_DTYPE_BF16 or residual.shape != y.shape:
                raise RuntimeError("deepipr_b200: fused residual join needs a bf16 block that keeps z and a residual "
                                   "of the output's shape")
            res = to_nhwc_bf16(residual.detach())
        L.check(L.load().pp_conv_block_fwd_res(
            C.byref(d), L.ptr(xc), L.ptr(prepared.wf), L.ptr(g), L.ptr(b), L.ptr(o.running_mean),
            L.ptr(o.running_var), L.ptr(z), L.ptr(y), L.ptr(save_mean), L.ptr(save_invstd), L.ptr(res), L.ptr(ws),
            C.c_size_t(nbytes), _stream()), "pp_conv_block_fwd")
        ctx.res_dtype = None if residual is None else residual.dtype
        ctx.link = o.link
        if need_grad:
            ctx.save_for_backward(xc, z, g, b, save_mean, save_invstd)
            # gradients that can be accumulated by the kernels themselves into a flat gradient buffer
            ctx.slots = [None, None, None]
            if o.direct_grad_ok:
                for k, t in enumerate((weight, gamma, beta)):
                    if t is not None and ctx.needs_input_grad[1 + k] and t.dtype == torch.float32:
                        slot = _flat_slot(t)
                        if slot is not None and (k == 0 or t.numel() == spec.O):
                            ctx.slots[k] = slot
                            slot[0].direct_begin(slot[1])
            ctx.prepared = prepared
            ctx.o = o
            ctx.x_dtype = x.dtype
            ctx.xshape = (N, Cx, H, W)
            ctx.wshape = weight.shape
            ctx.gshape = None if gamma is None else gamma.shape
            ctx.bshape = None if beta is None else beta.shape
            ctx.gdtype = None if gamma is None else gamma.dtype
            ctx.bdtype = None if beta is None else beta.dtype
        out_dtype = o.out_dtype or x.dtype
        if out_dtype != adt:
            y = y.to(out_dtype)
        # The output keeps the memory format of the input: channels_last in -> channels_last out (no copy; what the
        # nets of this package and autocast pipelines use), NCHW-contiguous in -> NCHW-contiguous out, because the
        # reference's model files call .view() on block outputs (models/alexnet_passport.py: x.view(x.size(0), -1),
        # passportconv2d.py:101 passport_candidates.view(b * c, h, w)) and a channels_last tensor is not viewable.
        if not x.is_contiguous(memory_format=torch.channels_last):
            y = y.contiguous()
        return y

    @staticmethod
    def backward(ctx, gy):
        xc, z, g, b, save_mean, save_invstd = ctx.saved_tensors
        o, spec = ctx.o, ctx.o.spec
        N, Cx, H, W = ctx.xshape
        dev = gy.device
        adt = act_dtype(o.dtype)
        gyc = to_nhwc(gy, o.dtype)
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        sw, sg, sb = ctx.slots
        link, role = ctx.link if ctx.link is not None else (None, None)
        dx_add = None
        if role == 'add' and link.g is not None:          # the residual path's gradient, parked by the closing block
            dx_add, link.g = link.g, None
        flags = (L.PP_FLAG_ACC_DW if sw else 0) | (L.PP_FLAG_ACC_DGAMMA if sg else 0) | (L.PP_FLAG_ACC_DBETA if sb else 0)
        d, key = make_desc(spec, N, H, W, o.norm, o.relu, o.z_f32, o.eps, o.momentum, o.algo, o.groups, flags, o.dtype)
        dx = torch.empty((N, Cx, H, W), dtype=adt, device=dev, memory_format=torch.channels_last) \
            if need_dx else None
        if sw:
            dw = sw[0].grad_view(sw[1])
        else:
            dw = torch.empty(ctx.wshape, dtype=torch.float32, device=dev) if need_dw else None
        dgamma = sg[0].grad_view(sg[1]).view(-1) if sg else torch.empty(spec.O, dtype=torch.float32, device=dev)
        dbeta = sb[0].grad_view(sb[1]).view(-1) if sb else torch.empty(spec.O, dtype=torch.float32, device=dev)
        if need_dx and ctx.prepared.wd is None:
            raise RuntimeError("deepipr_b200: dgrad weights were not prepared")
        ws, nbytes = workspace(d, key, L.PP_WS_BWD, dev)
        if OVERLAP_WGRAD and sw and need_dw and o.norm != L.PP_NORM_GN:
            # dgamma / dbeta / dz / dx here; the weight gradient on the side stream, straight into the flat buffer
            if _side_for(dev).pending:      # a weight gradient is in flight: share the SMs with it (PP_FLAG_SHARE_SM)
                d, key = make_desc(spec, N, H, W, o.norm, o.relu, o.z_f32, o.eps, o.momentum, o.algo, o.groups,
                                   flags | L.PP_FLAG_SHARE_SM, o.dtype)
            dzbuf = torch.empty((N,) + spec.out_hw(H, W) + (spec.O,), dtype=adt, device=dev)
            fused_add = None
            if dx_add is not None and dx is not None and adt == torch.bfloat16 and dx_add.shape == dx.shape:
                fused_add, dx_add = to_nhwc_bf16(dx_add), None          # summed in the dgrad epilogue
            L.check(L.load().pp_conv_block_bwd_dz(
                C.byref(d), L.ptr(gyc), L.ptr(ctx.prepared.wd), L.ptr(z), L.ptr(g), L.ptr(b), L.ptr(save_mean),
                L.ptr(save_invstd), L.ptr(dx), L.ptr(fused_add), L.ptr(dgamma), L.ptr(dbeta), L.ptr(dzbuf), L.ptr(ws),
                C.c_size_t(nbytes), _stream()), "pp_conv_block_bwd_dz")
            side = _side_for(dev)
            ws2, nbytes2 = workspace(d, key, L.PP_WS_BWD, dev, slot=1)
            side.stream.wait_stream(torch.cuda.current_stream(dev))
            L.check(L.load().pp_conv_wgrad(C.byref(d), L.ptr(dzbuf), L.ptr(xc), L.ptr(dw), L.ptr(ws2),
                                           C.c_size_t(nbytes2), C.c_void_p(side.stream.cuda_stream)), "pp_conv_wgrad")
            side.pending = True
            side.keep.append((dzbuf, xc, dw))
            if not side.cb_queued:           # whoever reads gradients after this backward pass finds them complete
                side.cb_queued = True
                torch.autograd.Variable._execution_engine.queue_callback(lambda: join_side(dev))
        else:
            L.check(L.load().pp_conv_block_bwd(
                C.byref(d), L.ptr(gyc), L.ptr(xc), L.ptr(ctx.prepared.wd), L.ptr(z), L.ptr(g), L.ptr(b),
                L.ptr(save_mean), L.ptr(save_invstd), L.ptr(dx), L.ptr(dw), L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws),
                C.c_size_t(nbytes), _stream()), "pp_conv_block_bwd")
        if dx_add is not None:                 # not summed by the kernel (other path / no dx of our own): one torch add
            dx = dx_add.to(adt) if dx is None else dx + dx_add.to(dx.dtype)
        if dx is not None and ctx.x_dtype != adt:
            dx = dx.to(ctx.x_dtype)
        gg = dgamma.reshape(ctx.gshape).to(ctx.gdtype) if (ctx.gshape is not None and ctx.needs_input_grad[2]) else None
        gb = dbeta.reshape(ctx.bshape).to(ctx.bdtype) if (ctx.bshape is not None and ctx.needs_input_grad[3]) else None
        # directly accumulated gradients are already in the flat buffer: autograd gets None for them
        for slot in (sw, sg, sb):
            if slot:
                slot[0].direct_done(slot[1])
        # the join is y = block + residual: the residual receives the upstream gradient unchanged (no kernel)
        g_res = gy.to(ctx.res_dtype) if (ctx.res_dtype is not None and ctx.needs_input_grad[6]) else None
        if g_res is not None and role == 'stash' and gy.dtype == torch.bfloat16 and ctx.res_dtype == torch.bfloat16:
            link.g, g_res = gyc, None          # the opening block of the unit adds it to its data gradient
        return dx, (None if sw else dw), (None if sg else gg), (None if sb else gb), None, None, g_res


@dataclass
class PassportCtx:
    """The passport of a block for the fused operator: pooled keys + signature."""
    S_skey: torch.Tensor
    S_key: torch.Tensor
    b: Optional[torch.Tensor]          # fp32 [O] (+-1) or None (no sign loss fed)
    alpha: float


class _PassportConvFn(torch.autograd.Function):
    """The passport block as ONE operator: gamma / beta from the passport, SignLoss.add, conv, batch-norm, affine, ReLU
    (pp_passport_conv_fwd; one cooperative kernel where the output tiles fit in tensor memory) and its backward
    (pp_passport_conv_bwd).  Reference: passportconv2d.py:142-175 + 209-223, passportconv2d_private.py:139-219,
    sign_loss.py:32-54."""

    @staticmethod
    def forward(ctx, x, weight, prepared: PreparedWeight, o: BlockOpts, pc: PassportCtx):
        spec = o.spec
        N, Cx, H, W = x.shape
        if Cx != spec.C:
            raise RuntimeError(f"input has {Cx} channels, block expects {spec.C}")
        P, Q = spec.out_hw(H, W)
        dev = x.device
        if prepared.dtype != o.dtype:
            raise RuntimeError("deepipr_b200: weight operands were prepared for another arithmetic type")
        adt = act_dtype(o.dtype)
        xc = to_nhwc(x.detach(), o.dtype)
        w = master_weight(weight)
        need_grad = any(ctx.needs_input_grad[:2])
        keep_z = need_grad or o.norm in (L.PP_NORM_BN_TRAIN, L.PP_NORM_GN)
        d, key = make_desc(spec, N, H, W, o.norm, o.relu, o.z_f32, o.eps, o.momentum, o.algo, o.groups, dtype=o.dtype)
        y = torch.empty((N, spec.O, P, Q), dtype=adt, device=dev, memory_format=torch.channels_last)
        z = torch.empty((N, P, Q, spec.O), dtype=torch.float32 if o.z_f32 else torch.bfloat16, device=dev) \
            if keep_z else None
        nstat = N * o.groups if o.norm == L.PP_NORM_GN else spec.O
        save_mean = torch.empty(nstat, dtype=torch.float32, device=dev)
        save_invstd = torch.empty(nstat, dtype=torch.float32, device=dev)
        gamma = torch.empty(spec.O, dtype=torch.float32, device=dev)
        beta = torch.empty(spec.O, dtype=torch.float32, device=dev)
        has_b = pc.b is not None
        stats = torch.zeros(2, dtype=torch.float32, device=dev) if has_b else None      # [sign loss, sign acc]
        ws, nbytes = workspace(d, key, L.PP_WS_FWD, dev)
        L.check(L.load().pp_passport_conv_fwd(
            C.byref(d), L.ptr(xc), L.ptr(prepared.wf), L.ptr(w), L.ptr(pc.S_skey), L.ptr(pc.S_key), None, None,
            L.ptr(pc.b), float(pc.alpha), L.ptr(o.running_mean), L.ptr(o.running_var), L.ptr(y), L.ptr(z),
            L.ptr(gamma), L.ptr(beta), L.ptr(save_mean), L.ptr(save_invstd),
            None if stats is None else C.c_void_p(stats.data_ptr()),
            None if stats is None else C.c_void_p(stats.data_ptr() + 4),
            L.ptr(ws), C.c_size_t(nbytes), _stream()), "pp_passport_conv_fwd")
        if need_grad:
            ctx.save_for_backward(xc, z, gamma, beta, save_mean, save_invstd)
            ctx.prepared, ctx.o, ctx.pc = prepared, o, pc
            ctx.x_dtype = x.dtype
            ctx.xshape = (N, Cx, H, W)
            ctx.wshape = weight.shape
        out_dtype = o.out_dtype or x.dtype
        if out_dtype != adt:
            y = y.to(out_dtype)
        if not x.is_contiguous(memory_format=torch.channels_last):
            y = y.contiguous()          # keep the input's memory format (see _ConvBlockFn.forward)
        ctx.mark_non_differentiable(gamma, beta)
        if has_b:
            loss, acc = stats[0], stats[1]
            ctx.mark_non_differentiable(acc)
            return y, gamma, beta, loss, acc
        return y, gamma, beta

    @staticmethod
    def backward(ctx, gy, g_gamma=None, g_beta=None, g_loss=None, g_acc=None):
        xc, z, gamma, beta, save_mean, save_invstd = ctx.saved_tensors
        o, spec, pc = ctx.o, ctx.o.spec, ctx.pc
        N, Cx, H, W = ctx.xshape
        dev = xc.device
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        adt = act_dtype(o.dtype)
        if gy is None:          # only the sign loss was used downstream
            gy = torch.zeros((N, spec.O) + spec.out_hw(H, W), dtype=adt, device=dev)
        gyc = to_nhwc(gy, o.dtype)
        d, key = make_desc(spec, N, H, W, o.norm, o.relu, o.z_f32, o.eps, o.momentum, o.algo, o.groups, 0, o.dtype)
        dx = torch.empty((N, Cx, H, W), dtype=adt, device=dev, memory_format=torch.channels_last) \
            if need_dx else None
        dw = torch.empty(ctx.wshape, dtype=torch.float32, device=dev) if need_dw else None
        dgamma = torch.empty(spec.O, dtype=torch.float32, device=dev)
        dbeta = torch.empty(spec.O, dtype=torch.float32, device=dev)
        gl = None if g_loss is None else g_loss.contiguous().float()
        if need_dx and ctx.prepared.wd is None:
            raise RuntimeError("deepipr_b200: dgrad weights were not prepared")
        ws, nbytes = workspace(d, key, L.PP_WS_BWD, dev)
        if need_dw:
            L.check(L.load().pp_passport_conv_bwd(
                C.byref(d), L.ptr(gyc), L.ptr(xc), L.ptr(ctx.prepared.wd), L.ptr(z), L.ptr(gamma), L.ptr(beta),
                L.ptr(save_mean), L.ptr(save_invstd), L.ptr(pc.S_skey), L.ptr(pc.S_key), L.ptr(pc.b), float(pc.alpha),
                L.ptr(gl), L.ptr(dx), L.ptr(dw), L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws), C.c_size_t(nbytes),
                _stream()), "pp_passport_conv_bwd")
        else:                   # frozen weight: plain block backward for dx
            L.check(L.load().pp_conv_block_bwd(
                C.byref(d), L.ptr(gyc), L.ptr(xc), L.ptr(ctx.prepared.wd), L.ptr(z), L.ptr(gamma), L.ptr(beta),
                L.ptr(save_mean), L.ptr(save_invstd), L.ptr(dx), None, L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws),
                C.c_size_t(nbytes), _stream()), "pp_conv_block_bwd")
        if dx is not None and ctx.x_dtype != adt:
            dx = dx.to(ctx.x_dtype)
        return dx, dw, None, None, None


def passport_conv(x, weight, prepared: PreparedWeight, opts: BlockOpts, pc: PassportCtx):
    """Returns (y, gamma[O], beta[O], sign_loss or None, sign_acc or None)."""
    require_cuda(x, "block input")
    require_cuda(weight, "conv weight")
    out = _PassportConvFn.apply(x, weight, prepared, opts, pc)
    if len(out) == 3:
        return out[0], out[1], out[2], None, None
    return out


def conv_block(x, weight, gamma, beta, prepared: PreparedWeight, opts: BlockOpts, residual=None):
    """``residual`` (optional, >= 0 everywhere, e.g. another block's post-ReLU output): the residual join of a basic
    unit folded into the block's last pass, y = block(x) + residual."""
    require_cuda(x, "block input")
    require_cuda(weight, "conv weight")
    if residual is not None:
        require_cuda(residual, "residual")
    return _ConvBlockFn.apply(x, weight, gamma, beta, prepared, opts, residual)


class _AddReluFn(torch.autograd.Function):
    """y = relu(a + b) for the residual join of a basic block (resnet_passport_private.py:78-85)."""

    @staticmethod
    def forward(ctx, a, b):
        y = torch.empty_like(a)
        L.check(L.load().pp_add_relu_fwd(C.c_size_t(a.numel()), L.ptr(a), L.ptr(b), L.ptr(y), _stream()),
                "pp_add_relu_fwd")
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        if gy.stride() != y.stride() or gy.dtype != y.dtype:
            gy = gy.to(y.dtype).contiguous(memory_format=torch.channels_last) if y.dim() == 4 and \
                y.is_contiguous(memory_format=torch.channels_last) else gy.to(y.dtype).contiguous()
        gx = torch.empty_like(y)
        L.check(L.load().pp_add_relu_bwd(C.c_size_t(y.numel()), L.ptr(gy), L.ptr(y), L.ptr(gx), _stream()),
                "pp_add_relu_bwd")
        return gx, gx


def add_relu(a, b):
    """relu(a + b) on the GPU: fused kernel for dense bf16 tensors of identical layout (the training configuration);
    other dtypes / layouts use the stock CUDA ops.  CPU tensors are rejected like everywhere else in this package."""
    require_cuda(a, "residual input")
    if (a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.shape == b.shape
            and a.stride() == b.stride()
            and (a.is_contiguous() or (a.dim() == 4 and a.is_contiguous(memory_format=torch.channels_last)))):
        return _AddReluFn.apply(a, b)
    return torch.relu(a + b)


class _MaxPoolFn(torch.autograd.Function):
    """nn.MaxPool2d(k, s, p) on a dense channels_last tensor (pp_maxpool_fwd / _bwd): the pooling layers of the
    reference nets (resnet_passport*.py ImageNet stem, alexnet_passport.py:37-38)."""

    @staticmethod
    def forward(ctx, x, k, s, p):
        N, Cc, H, W = x.shape
        P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        xc = x.detach().contiguous(memory_format=torch.channels_last)
        f32 = int(xc.dtype == torch.float32)
        y = torch.empty((N, Cc, P, Q), dtype=xc.dtype, device=x.device, memory_format=torch.channels_last)
        idx = torch.empty((N, P, Q, Cc), dtype=torch.uint8, device=x.device)
        L.check(L.load().pp_maxpool_fwd(N, H, W, Cc, k, s, p, L.ptr(xc), f32, L.ptr(y), L.ptr(idx), _stream()),
                "pp_maxpool_fwd")
        ctx.save_for_backward(idx)
        ctx.geom = (N, Cc, H, W, k, s, p)
        return y

    @staticmethod
    def backward(ctx, gy):
        (idx,) = ctx.saved_tensors
        N, Cc, H, W, k, s, p = ctx.geom
        g = gy.contiguous(memory_format=torch.channels_last)
        f32 = int(g.dtype == torch.float32)
        if g.dtype not in (torch.float32, torch.bfloat16):
            g, f32 = g.float(), 1
        dx = torch.empty((N, Cc, H, W), dtype=g.dtype, device=g.device, memory_format=torch.channels_last)
        L.check(L.load().pp_maxpool_bwd(N, H, W, Cc, k, s, p, L.ptr(g), L.ptr(idx), f32, L.ptr(dx), _stream()),
                "pp_maxpool_bwd")
        return dx.to(gy.dtype), None, None, None


def max_pool2d_supported(x, k, s, p):
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.bfloat16, torch.float32) and x.shape[1] % 8 == 0
            and x.is_contiguous(memory_format=torch.channels_last) and 1 <= k <= 15 and 2 * p <= k
            and x.shape[2] + 2 * p >= k and x.shape[3] + 2 * p >= k)


def max_pool2d(x, k, s, p):
    """F.max_pool2d(x, k, s, p) for dense channels_last bf16 / fp32 CUDA tensors with C % 8 == 0."""
    require_cuda(x, "max-pool input")
    if not max_pool2d_supported(x, k, s, p):
        raise RuntimeError("deepipr_b200.max_pool2d: unsupported input (needs channels_last bf16/fp32, C % 8 == 0)")
    return _MaxPoolFn.apply(x, int(k), int(s), int(p))


# ---------------------------------------------------------------- raw building blocks (tests / profiling)
def conv_fwd_raw(x, prepared: PreparedWeight, spec: ConvSpec, z_f32=False, algo=None):
    require_cuda(x, "input")
    N, _, H, W = x.shape
    P, Q = spec.out_hw(H, W)
    tf32 = prepared.dtype == L.PP_DTYPE_TF32
    z_f32 = bool(z_f32) or tf32
    xc = to_nhwc(x, prepared.dtype)
    d, key = make_desc(spec, N, H, W, z_f32=int(z_f32), algo=algo, dtype=prepared.dtype)
    z = torch.empty((N, P, Q, spec.O), dtype=torch.float32 if z_f32 else torch.bfloat16, device=x.device)
    ws, nbytes = workspace(d, key, L.PP_WS_FWD, x.device)
    L.check(L.load().pp_conv_fwd_raw(C.byref(d), L.ptr(xc), L.ptr(prepared.wf), L.ptr(z), L.ptr(ws), C.c_size_t(nbytes),
                                     _stream()), "pp_conv_fwd_raw")
    return z  # NHWC


def conv_dgrad(dz_nhwc, prepared: PreparedWeight, spec: ConvSpec, N, H, W, algo=None):
    require_cuda(dz_nhwc, "dz")
    d, key = make_desc(spec, N, H, W, algo=algo, dtype=prepared.dtype)
    dz = dz_nhwc.to(act_dtype(prepared.dtype)).contiguous()
    dx = torch.empty((N, H, W, spec.C), dtype=act_dtype(prepared.dtype), device=dz.device)
    L.check(L.load().pp_conv_dgrad(C.byref(d), L.ptr(dz), L.ptr(prepared.wd), L.ptr(dx), _stream()), "pp_conv_dgrad")
    return dx  # NHWC


def conv_wgrad(dz_nhwc, x, spec: ConvSpec, algo=None, dtype=L.PP_DTYPE_BF16):
    require_cuda(dz_nhwc, "dz")
    N, _, H, W = x.shape
    xc = to_nhwc(x, dtype)
    d, key = make_desc(spec, N, H, W, algo=algo, dtype=dtype)
    dz = dz_nhwc.to(act_dtype(dtype)).contiguous()
    dw = torch.empty((spec.O, spec.C, spec.kh, spec.kw), dtype=torch.float32, device=dz.device)
    ws, nbytes = workspace(d, key, L.PP_WS_BWD, dz.device)
    L.check(L.load().pp_conv_wgrad(C.byref(d), L.ptr(dz), L.ptr(xc), L.ptr(dw), L.ptr(ws), C.c_size_t(nbytes),
                                   _stream()), "pp_conv_wgrad")
    return dw

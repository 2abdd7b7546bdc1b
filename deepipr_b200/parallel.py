"""Data-parallel plumbing for the passport training loop: one process per GPU, gradients all-reduced over
NCCL (NVLink 5 / NVSwitch) on flat fp32 buckets only — no activation or buffer traffic per step.

The reference's only multi-GPU mechanism is nn.DataParallel (experiments/trainer.py:92-93,
experiments/trainer_private.py:110-111), which replicates the module every forward and silently drops the
SignLoss of the replicas.  Here every rank owns a full replica; the passport-derived
gamma/beta and the sign loss depend on the weights and keys only, so they are computed redundantly and
identically on every rank and their gradient survives the averaging unchanged.  BatchNorm statistics stay
per-rank (what DataParallel does too).

  FlatParams      parameters and gradients re-homed as views of two flat fp32 buffers
  GradBuckets     bucketed (size-bounded, reverse order) asynchronous all-reduce launched from
                  post-accumulate-grad hooks so communication overlaps the rest of backward
  FlatSGD         torch.optim.Optimizer whose step is ONE pp_sgd_step launch per param group on the flat buffers
                  (momentum, weight decay: experiments/classification.py:47-50)
"""
import ctypes as C
import weakref

import os

import torch
import torch.distributed as dist

from . import _lib as L
from . import functional as F_


class FlatParams:
    """Re-home `params` (fp32, same device) into one flat buffer; same for their gradients."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        seen, uniq = set(), []
        for p in self.params:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("FlatParams needs all parameters on one device with one dtype")
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 63) // 64 * 64          # keep every view 256-byte aligned
        self.numel = off
        self.flat = torch.zeros(off, dtype=dt, device=dev)
        self.flat_grad = torch.zeros(off, dtype=dt, device=dev)
        #: Direct accumulation: the fused conv operator adds the weight / BN-affine gradients of ConvBlocks into
        #: flat_grad from inside its own kernels (PP_FLAG_ACC_*) instead of handing them to autograd, which would
        #: spend one `grad += g` launch per parameter and pass (~120 tiny launches per V2 step).  Set False to get
        #: the plain autograd path back (identical results; tests compare the two).
        self.direct = True
        self._direct_pending = [0] * len(self.params)   # forward uses of parameter i whose backward has not run yet
        self._direct_uses = [0] * len(self.params)      # direct uses of parameter i since the last zero_grad()
        self.on_ready = None                             # GradBuckets: called with i when parameter i's grad is final
        with torch.no_grad():
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                view = self.flat[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
                p._pp_flat_slot = (weakref.ref(self), i)
        F_.bump_weight_epoch()

    # ---- direct accumulation bookkeeping (functional._ConvBlockFn)
    def direct_begin(self, i):
        self._direct_pending[i] += 1
        self._direct_uses[i] += 1

    def is_direct(self, i):
        """True when parameter i's gradient is being produced by direct accumulation in the current step: its
        readiness is reported by direct_done(), and the AccumulateGrad hook autograd still fires for it (with an
        undefined gradient) must not be counted a second time."""
        return self._direct_uses[i] > 0

    def direct_done(self, i):
        self._direct_pending[i] -= 1
        if self._direct_pending[i] == 0 and self.on_ready is not None:
            self.on_ready(i)

    def grad_view(self, i):
        p, o = self.params[i], self.offsets[i]
        return self.flat_grad[o:o + p.numel()].view_as(p)

    def ensure_grad_views(self):
        """After an external zero_grad(set_to_none=True) the views are gone: restore them (copying any grad in)."""
        for i, p in enumerate(self.params):
            view = self.grad_view(i)
            if p.grad is None:
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view

    def zero_grad(self):
        if self.flat_grad.is_cuda:
            F_.join_side(self.flat_grad.device)
        self.flat_grad.zero_()
        self.ensure_grad_views()
        # forwards whose backward never ran (evaluation with grad enabled, an aborted step) must not leak into the
        # next step's readiness count
        self._direct_pending = [0] * len(self.params)
        self._direct_uses = [0] * len(self.params)


class GradBuckets:
    """Size-bounded buckets over FlatParams.flat_grad, all-reduced (mean) asynchronously as they fill."""

    def __init__(self, flat: FlatParams, bucket_bytes=None, process_group=None, overlap=True):
        if bucket_bytes is None:      # PP_BUCKET_MB: A/B runs of the bucket size without touching the callers
            bucket_bytes = int(float(os.environ.get("PP_BUCKET_MB", "25")) * (1 << 20))
        self.flat = flat
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.overlap = overlap
        self._avg_op = None
        if dist.is_initialized() and dist.get_backend(process_group) == "nccl":
            self._avg_op = dist.ReduceOp.AVG
        # buckets in reverse parameter order (gradients become ready roughly back to front)
        self.buckets = []            # (start, end) element ranges of flat_grad
        self.bucket_of = [0] * len(flat.params)
        cur_end = flat.numel
        cur_start = cur_end
        members = []
        elem_bytes = flat.flat_grad.element_size()
        for i in reversed(range(len(flat.params))):
            start = flat.offsets[i]
            if members and (cur_end - start) * elem_bytes > bucket_bytes:
                self._close(cur_start, cur_end, members)
                cur_end, members = cur_start, []
            cur_start = start
            members.append(i)
        if members:
            self._close(cur_start, cur_end, members)
        self._pending = [len(m) for (_, _, m) in self.buckets]
        self._next = 0               # first bucket not yet handed to the communicator
        self._works = []
        self._hooks = []
        self._ready = [False] * len(flat.params)   # parameter i reported its gradient in the current step
        self._accumulate = False     # no_sync(): gradients accumulate locally, nothing is counted or reduced
        if overlap and self.world > 1:
            for i, p in enumerate(flat.params):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))
            # parameters whose gradients are accumulated directly by the kernels never reach AccumulateGrad:
            # FlatParams tells us when their last pending backward has run
            flat.on_ready = self._param_ready

    def _close(self, start, end, members):
        idx = len(self.buckets)
        for i in members:
            self.bucket_of[i] = idx
        self.buckets.append((start, end, list(members)))

    def no_sync(self):
        """Context manager for gradient accumulation: backward passes inside it only add into the flat gradient
        buffer (no bucket is counted or all-reduced).  The first backward OUTSIDE it, followed by finish(), reduces
        the accumulated sums — the contract of torch DDP.no_sync()."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            prev, self._accumulate = self._accumulate, True
            try:
                yield self
            finally:
                self._accumulate = prev
        return ctx()

    def _param_ready(self, i):
        if self._accumulate:
            return
        if self._ready[i]:
            # A second backward before finish() (gradient accumulation, retain_graph): this parameter's bucket may
            # already be divided by the world size and in flight, so adding an un-reduced local gradient on top would
            # make the ranks diverge silently.  Refuse loudly instead.
            raise RuntimeError(
                "deepipr_b200.GradBuckets: a parameter reported a second gradient before finish(); run all but the "
                "last backward of a step inside GradBuckets.no_sync() and call finish() once per optimizer step")
        self._ready[i] = True
        b = self.bucket_of[i]
        self._pending[b] -= 1
        # collectives must be issued in the same order on every rank: launch strictly in bucket order
        while self._next < len(self.buckets) and self._pending[self._next] == 0:
            self._launch(self._next)
            self._next += 1

    def _make_hook(self, i):
        def hook(p):
            if self._accumulate:
                view = self.flat.grad_view(i)
                if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                    view.add_(p.grad)
                    p.grad = view
                return
            if self.flat.is_direct(i):
                return          # FlatParams.direct_done() reports this parameter (see FlatParams.is_direct)
            view = self.flat.grad_view(i)
            if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view
            self._param_ready(i)
        return hook

    def _launch(self, b):
        start, end, _ = self.buckets[b]
        chunk = self.flat.flat_grad[start:end]
        F_.join_side(chunk.device if chunk.is_cuda else None)   # weight gradients still in flight on the side stream
        if self.world > 1:
            if self._avg_op is not None:     # NCCL: the mean is taken inside the collective (no pre-scaling launch)
                self._works.append(dist.all_reduce(chunk, op=self._avg_op, group=self.group, async_op=True))
            else:
                chunk.div_(self.world)       # gloo has no AVG: SUM of pre-divided gradients == mean
                self._works.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Call after the last backward() of a step: reduce whatever the hooks have not launched yet and wait."""
        if self.world > 1:
            # a parameter that received no gradient this step (p.grad is None after an external
            # zero_grad(set_to_none=True)) must contribute zeros, not last step's values still sitting in the flat buffer
            for i, p in enumerate(self.flat.params):
                if p.grad is None:
                    self.flat.grad_view(i).zero_()
            self.flat.ensure_grad_views()
            while self._next < len(self.buckets):   # no-overlap mode, or parameters that got no gradient this step
                self._launch(self._next)
                self._next += 1
            for w in self._works:
                w.wait()
        self._works = []
        self._next = 0
        self._pending = [len(m) for (_, _, m) in self.buckets]
        self._ready = [False] * len(self.flat.params)

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if self.flat.on_ready == self._param_ready:
            self.flat.on_ready = None


def broadcast_state(module, src=0, group=None):
    """Make every rank start from rank `src`'s parameters and buffers (keys, signatures, BN statistics)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + [b for b in module.buffers() if b is not None]:
            dist.broadcast(t, src=src, group=group)
    for m in module.modules():
        if hasattr(m, 'invalidate_cache'):
            m.invalidate_cache()


class FlatSGD(torch.optim.Optimizer):
    """SGD(momentum, weight_decay) over FlatParams: one fused kernel launch per step (pp_sgd_step).

    Differences from torch.optim.SGD, by construction of the flat buffer: a parameter that received no gradient is
    treated as having a zero gradient (weight decay and momentum still apply to it; torch skips it), and the update
    also runs over the zero padding between parameters (which stays zero)."""

    def __init__(self, flat: FlatParams, lr=0.01, momentum=0.9, weight_decay=1e-4):
        self.flat = flat
        defaults = dict(lr=lr, momentum=momentum, weight_decay=weight_decay)
        super().__init__(flat.params, defaults)
        self._buf = torch.zeros_like(flat.flat)
        self._steps = 0
        # optional: hyper-parameters in device memory, so that a CUDA-graph replay of step() follows lr schedules
        self._hyper = None
        self._hyper_host = None

    def use_device_hyper(self, on=True):
        """step() reads {lr, momentum, weight_decay, first-step flag} from a device buffer (pp_sgd_step_dev)."""
        if on and self._hyper is None:
            self._hyper = torch.zeros(4, dtype=torch.float32, device=self.flat.flat.device)
            self._hyper_host = None
        if not on:
            self._hyper = None

    def sync_hyper(self):
        """Refresh the device copy of the hyper-parameters if they changed (16-byte copy; call before a replay)."""
        if self._hyper is None:
            return
        g = self.param_groups[0]
        vals = (float(g['lr']), float(g['momentum']), float(g['weight_decay']), float(self._steps == 0))
        if vals != self._hyper_host:
            self._hyper.copy_(torch.tensor(vals, dtype=torch.float32), non_blocking=False)
            self._hyper_host = vals

    def add_param_group(self, param_group):
        if getattr(self, "param_groups", None):
            raise ValueError("FlatSGD updates ONE flat buffer with one (lr, momentum, weight_decay): a second param "
                             "group is not supported (build a second FlatParams + FlatSGD instead)")
        super().add_param_group(param_group)

    def state_dict(self):
        """torch layout plus the flat momentum buffer and the step count (momentum lives outside Optimizer.state)."""
        sd = super().state_dict()
        sd["flat_momentum"] = self._buf.clone()
        sd["flat_steps"] = self._steps
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        buf, steps = state_dict.pop("flat_momentum", None), state_dict.pop("flat_steps", 0)
        super().load_state_dict(state_dict)
        if buf is not None:
            if buf.numel() != self._buf.numel():
                raise ValueError("FlatSGD.load_state_dict: momentum buffer size differs from this FlatParams")
            self._buf.copy_(buf.to(self._buf.device))
        self._steps = int(steps)

    def zero_grad(self, set_to_none=False):
        self.flat.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        F_.require_cuda(self.flat.flat, "FlatSGD parameters")
        F_.join_side(self.flat.flat.device)
        self.flat.ensure_grad_views()
        if len(self.param_groups) != 1:
            raise ValueError("FlatSGD supports exactly one param group")
        g = self.param_groups[0]
        if self._hyper is not None:
            if not torch.cuda.is_current_stream_capturing():
                self.sync_hyper()
            L.check(L.load().pp_sgd_step_dev(
                C.c_size_t(self.flat.numel), L.ptr(self.flat.flat), L.ptr(self.flat.flat_grad), L.ptr(self._buf),
                L.ptr(self._hyper), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pp_sgd_step_dev")
            self._steps += 1
            F_.bump_weight_epoch()
            return loss
        L.check(L.load().pp_sgd_step(
            C.c_size_t(self.flat.numel), L.ptr(self.flat.flat), L.ptr(self.flat.flat_grad), L.ptr(self._buf),
            float(g['lr']), float(g['momentum']), float(g['weight_decay']), int(self._steps == 0),
            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pp_sgd_step")
        self._steps += 1
        F_.bump_weight_epoch()   # parameters changed behind autograd's back: invalidate bf16 operand caches
        return loss

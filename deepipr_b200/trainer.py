"""Training / evaluation loops with the reference's step semantics and return dictionaries
(experiments/trainer.py:99-214 `Trainer`, experiments/trainer_private.py:37-257 `TrainerPrivate`/`TesterPrivate`),
written for one process per GPU: where the reference wraps the model in nn.DataParallel
(trainer.py:92-93, trainer_private.py:110-111) this module all-reduces flat gradient buckets over NCCL
(deepipr_b200.parallel).  Used by bench.py and the tests when the reference checkout is absent; with the
reference present its own trainers run unchanged on the patched layers (deepipr_b200.patch_reference).
"""
import time

import torch
import torch.nn.functional as F

from . import _lib as L_
from . import functional as F_
from .layers import PassportBlock, PassportPrivateBlock, SignLoss
from .parallel import GradBuckets


def accuracy(output, target, topk=(1,)):
    """precision@k in percent (trainer.py:28-43)."""
    with torch.no_grad():
        maxk = max(topk)
        _, pred = output.topk(maxk, 1, True, True)
        correct = pred.t().eq(target.view(1, -1))
        return [correct[:k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / target.size(0)) for k in topk]


def sign_loss_modules(model):
    return [m for m in model.modules() if isinstance(m, SignLoss)]


def test_signature(model):
    """Fraction of signature bits recovered per passport layer (trainer_private.py:37-71), all layers verified by
    ONE kernel launch and one device->host read (pp_signature_verify).  Same keys / values as the reference's dict;
    the bits are those of the per-layer get_scale() path (test_signature_per_layer)."""
    model.eval()
    names, entries, res = [], [], {}
    with torch.no_grad():
        for name, m in model.named_modules():
            if isinstance(m, PassportPrivateBlock):
                tag = 'private_' + name
            elif isinstance(m, PassportBlock):
                if m.scale is not None:      # get_scale() returns the learnable scale, not the passport one (:143-144)
                    res['public_' + name] = (m.get_scale().view(-1).sign() == m.b).float().mean().item()
                    continue
                tag = 'public_' + name
            else:
                continue
            m._check_conv()
            S_skey, _ = m._pooled_keys()
            names.append(tag)
            entries.append((m.weight, S_skey, m.b))
        matched, Os, _ = F_.signature_verify(entries)
        if entries:
            det = (matched.float() / torch.tensor(Os, dtype=torch.float32, device=matched.device)).tolist()
            res.update(zip(names, det))
    # the reference's dict is ordered by named_modules(); keep that order
    order = [('private_' if isinstance(m, PassportPrivateBlock) else 'public_') + n for n, m in model.named_modules()
             if isinstance(m, (PassportPrivateBlock, PassportBlock))]
    return {k: res[k] for k in order}


def test_signature_per_layer(model):
    """The reference's loop verbatim: one get_scale() + one .item() per passport layer."""
    model.eval()
    res = {}
    with torch.no_grad():
        for name, m in model.named_modules():
            if isinstance(m, PassportPrivateBlock):
                res['private_' + name] = (m.get_scale(ind=1).view(-1).sign() == m.b).float().mean().item()
            if isinstance(m, PassportBlock):
                res['public_' + name] = (m.get_scale().view(-1).sign() == m.b).float().mean().item()
    return res


class StepRunner:
    """One optimisation step, V1 (one forward) or V2/V3 (public + private forward, one backward).

    Nothing in here reads a value back to the host.  With ``fused_loss`` (default) the cross-entropy, the precision@1
    and the logits gradient of each pass come from one kernel (pp_ce_top1) instead of the reference loop's
    F.cross_entropy + accuracy() chains; after every step ``self.metrics`` holds the four numbers the reference loop
    reads with four ``.item()`` calls — [loss, sign_loss, acc of pass 0, acc of pass 1] — as ONE device tensor."""

    def __init__(self, model, optimizer, private, buckets: GradBuckets = None, autocast=True, fused_loss=True):
        self.model, self.optimizer, self.private, self.buckets = model, optimizer, private, buckets
        self.autocast = autocast
        self.fused_loss = fused_loss
        self._losses = sign_loss_modules(model)
        self.metrics = None

    def _loss_and_acc(self, pred, target):
        if self.fused_loss:
            return F_.ce_top1(pred, target)
        return F.cross_entropy(pred.float(), target), accuracy(pred, target)[0].reshape(())

    def forward_backward(self, data, target):
        """Returns (loss, sign_loss, [logits per pass]) as device tensors — no host sync in here."""
        self.optimizer.zero_grad()
        for m in self._losses:
            m.reset()
        preds, accs = [], []
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.autocast):
            loss = None
            for ind in range(2 if self.private else 1):            # "backprop to two graph at once"
                pred = self.model(data, ind=ind) if self.private else self.model(data)
                l, a = self._loss_and_acc(pred, target)
                loss = l if loss is None else loss + l
                preds.append(pred)
                accs.append(a)
        sign_loss = torch.zeros((), device=data.device)
        for m in self._losses:
            sign_loss = sign_loss + m.loss
        (loss + sign_loss).backward()
        if self.buckets is not None:
            self.buckets.finish()
        with torch.no_grad():
            self.metrics = torch.stack([loss.detach().float(), sign_loss.detach().float(), accs[0].float(),
                                        accs[-1].float()])
        return loss, sign_loss, preds

    def step(self, data, target):
        loss, sign_loss, preds = self.forward_backward(data, target)
        self.optimizer.step()
        return loss, sign_loss, preds


class GraphedStepRunner:
    """StepRunner.step (zero_grad, forward(s), loss, backward, fused SGD) captured ONCE into a CUDA graph and replayed.

    One V2 step issues ~650 kernel launches through Python, autograd and ctypes: ~10 ms of host time, which at the
    reference's own batch sizes (64: train_v1.py:15, 256: training.sh:4) is several times the GPU time.  A replay costs
    one launch.  What makes the step capturable: the library allocates nothing and never synchronises, TMA descriptors
    and kernel arguments are plain launch parameters, the operand / key-pool caches take the same branches every step,
    and the SGD hyper-parameters live in device memory (pp_sgd_step_dev), so a learning-rate schedule needs no
    re-capture.  Batch shape is fixed.  Under DDP the bucketed NCCL all-reduces that GradBuckets launches from the
    backward hooks are captured with the step (they fork from / join the capturing stream like any side-stream work):
    every rank replays the same graph, so the collectives stay matched; the warm-up steps create the communicator
    eagerly, and pending eager work is drained before the capture starts.

    The warm-up steps capture needs are rolled back (parameters, momentum, buffers), so a graphed run follows exactly
    the trajectory of the eager one."""

    def __init__(self, runner: StepRunner, data, target, warmup=2):
        from .parallel import FlatSGD
        if not isinstance(runner.optimizer, FlatSGD):
            raise RuntimeError("GraphedStepRunner needs parallel.FlatSGD (its update is one capturable launch)")
        F_.require_cuda(data, "graph input")
        self.runner, self.opt = runner, runner.optimizer
        self.x, self.t = data.detach().clone(), target.detach().clone()
        flat, opt, model = self.opt.flat, self.opt, runner.model
        opt.use_device_hyper(True)
        saved = (flat.flat.clone(), opt._buf.clone(), opt._steps, [b.clone() for b in model.buffers()])
        side = torch.cuda.Stream(device=data.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                runner.step(self.x, self.t)
        torch.cuda.current_stream().wait_stream(side)
        if runner.buckets is not None and runner.buckets.world > 1:
            import torch.distributed as dist
            torch.cuda.synchronize()            # no eager collective may still be in flight when the capture starts
            dist.barrier(group=runner.buckets.group)
            torch.cuda.synchronize()
        # everything the captured kernels point at that was allocated BEFORE the capture must outlive the graph
        self._keep = [m.__dict__.get('_pp_keypool') for m in model.modules()] + list(F_._workspace.values())
        self.graph = torch.cuda.CUDAGraph()
        lib = L_.load()
        before = int(lib.pp_launch_count(0))
        with torch.cuda.graph(self.graph):
            self.out = runner.step(self.x, self.t)
            self.metrics = runner.metrics
        #: kernels of libpassport_sm100 inside the captured step (a replay re-launches exactly these)
        self.launches_per_replay = int(lib.pp_launch_count(0)) - before
        with torch.no_grad():                              # roll the warm-up back
            flat.flat.copy_(saved[0])
            opt._buf.copy_(saved[1])
            opt._steps = saved[2]
            for b, v in zip(model.buffers(), saved[3]):
                b.copy_(v)
        F_.bump_weight_epoch()

    def matches(self, data, target):
        return (self.graph is not None and data.shape == self.x.shape and target.shape == self.t.shape
                and data.dtype == self.x.dtype)

    def release(self):
        """Destroy the captured graph (call before tearing down a process group whose collectives it captured: NCCL
        does not finish destroying a communicator while graphs that reference it are alive)."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph = None
            self.out = self.metrics = None
            self._keep = []

    def step(self, data, target):
        if data.data_ptr() != self.x.data_ptr():
            self.x.copy_(data, non_blocking=True)
        if target.data_ptr() != self.t.data_ptr():
            self.t.copy_(target, non_blocking=True)
        self.opt.sync_hyper()
        self.graph.replay()
        self.opt._steps += 1
        F_.bump_weight_epoch()
        self.runner.metrics = self.metrics
        return self.out


def _cat_trigger(data, target, wm_iter, wm_loader, device):
    """V3: append the next trigger-set minibatch (trainer_private.py:135-146)."""
    try:
        wm_data, wm_target = next(wm_iter)
    except StopIteration:
        wm_iter = iter(wm_loader)
        wm_data, wm_target = next(wm_iter)
    wm_data = wm_data.to(device, non_blocking=True)
    wm_target = wm_target.to(device, non_blocking=True)
    return torch.cat([data, wm_data], dim=0), torch.cat([target, wm_target], dim=0), wm_iter


class _TrainerBase:
    private = False

    def __init__(self, model, optimizer, scheduler, device, buckets=None, autocast=True, verbose=False,
                 use_graph=False):
        self.model, self.optimizer, self.scheduler, self.device = model, optimizer, scheduler, device
        self.runner = StepRunner(model, optimizer, self.private, buckets, autocast)
        self.verbose = verbose
        self.use_graph = use_graph
        self._graphed = None
        self._copy_stream = None
        self.log_every = 0          # k > 0: read the running metrics back every k batches (progress reporting)
        self.on_log = None

    def _step(self, data, target):
        """Eager step, or the CUDA-graph replay when enabled and the batch has the captured shape."""
        if not self.use_graph:
            return self.runner.step(data, target)
        if self._graphed is None:
            self._graphed = GraphedStepRunner(self.runner, data, target)
        if self._graphed.matches(data, target):
            return self._graphed.step(data, target)
        return self.runner.step(data, target)            # ragged last batch

    def _fetch(self, it):
        """Next (data, target) on the device.  Pinned host batches (DataLoader(pin_memory=True)) are copied on a side
        stream, so the host->device copy of batch i+1 overlaps the compute of batch i; anything else is moved in line
        like the reference loop does (trainer_private.py:149-151)."""
        batch = next(it, None)
        if batch is None:
            return None
        data, target = batch[0], batch[1]
        if (not data.is_cuda) and data.is_pinned():
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            with torch.cuda.stream(self._copy_stream):
                data = data.to(self.device, non_blocking=True)
                target = target.to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            return data, target, ev
        return data.to(self.device, non_blocking=True), target.to(self.device, non_blocking=True), None

    def _epoch(self, dataloader, wm_dataloader, log_every=0, on_log=None):
        """The minibatch loop of Trainer.train / TrainerPrivate.train (trainer.py:123-148, trainer_private.py:131-177).
        The reference reads loss / sign loss / accuracies back with four .item() calls after every batch; here they
        are summed on the device and read ONCE after the last batch — or, with log_every=k, as one 16-byte copy every
        k batches (what a progress line needs).  Returns ([loss, sign_loss, acc0, acc1] sums, #batches)."""
        self.model.train()
        wm_iter = iter(wm_dataloader) if wm_dataloader is not None else None
        meters = torch.zeros(4, device=self.device)
        n = 0
        it = iter(dataloader)
        nxt = self._fetch(it)
        while nxt is not None:
            data, target, ev = nxt
            if ev is not None:
                cur = torch.cuda.current_stream()
                cur.wait_event(ev)
                data.record_stream(cur)
                target.record_stream(cur)
            nxt = self._fetch(it)
            if wm_iter is not None:
                data, target, wm_iter = _cat_trigger(data, target, wm_iter, wm_dataloader, self.device)
            self._step(data, target)
            meters += self.runner.metrics
            n += 1
            if log_every and n % log_every == 0:
                vals = meters.tolist()                      # the one device->host read of this step
                if on_log is not None:
                    on_log(n, vals)
        return meters.tolist(), max(n, 1)

    def _sign_acc(self):
        accs = [m.acc for m in sign_loss_modules(self.model)]
        if not accs:
            return 0.0
        total = torch.zeros((), device=self.device)
        for a in accs:
            total = total + a
        return (total / len(accs)).item()


class Trainer(_TrainerBase):
    """V1 / baseline loop (trainer.py:99-180): returns loss, sign_loss (mean per batch), sign_acc, acc, time."""

    def train(self, e, dataloader, wm_dataloader=None):
        t0 = time.time()
        (loss_m, sign_m, acc_m, _), n = self._epoch(dataloader, wm_dataloader, self.log_every, self.on_log)
        if self.scheduler is not None:
            self.scheduler.step()
        return {'loss': loss_m / n, 'sign_loss': sign_m / n, 'sign_acc': self._sign_acc(), 'acc': acc_m / n,
                'time': time.time() - t0}

    def test(self, dataloader, msg='Testing Result'):
        self.model.eval()
        loss_m = acc_m = cnt = 0
        t0 = time.time()
        with torch.no_grad():
            for load in dataloader:
                data, target = load[:2]
                data = data.to(self.device, non_blocking=True)
                target = target.to(self.device, non_blocking=True)
                pred = self.model(data).float()
                loss_m += F.cross_entropy(pred, target, reduction='sum').item()
                acc_m += pred.argmax(1).eq(target).sum().item()
                cnt += data.size(0)
        return {'loss': loss_m / cnt, 'acc': 100 * acc_m / cnt, 'time': time.time() - t0}


class TrainerPrivate(_TrainerBase):
    """V2 / V3 loop (trainer_private.py:118-211): sign_loss is the SUM over batches, as in the reference."""
    private = True

    def train(self, e, dataloader, wm_dataloader=None):
        t0 = time.time()
        (loss_m, sign_m, pub_m, priv_m), n = self._epoch(dataloader, wm_dataloader, self.log_every, self.on_log)
        if self.scheduler is not None:
            self.scheduler.step()
        return {'loss': loss_m / n, 'sign_loss': sign_m, 'sign_acc': self._sign_acc(), 'acc_public': pub_m / n,
                'acc_private': priv_m / n, 'time': time.time() - t0}

    def test(self, dataloader, msg='Testing Result'):
        self.model.eval()
        out = {}
        for ind, key in enumerate(('public', 'private')):
            loss_m = acc_m = cnt = 0
            t0 = time.time()
            with torch.no_grad():
                for load in dataloader:
                    data, target = load[:2]
                    data = data.to(self.device, non_blocking=True)
                    target = target.to(self.device, non_blocking=True)
                    pred = self.model(data, ind=ind).float()
                    loss_m += F.cross_entropy(pred, target, reduction='sum').item()
                    acc_m += pred.argmax(1).eq(target).sum().item()
                    cnt += data.size(0)
            out.update({'loss_' + key: loss_m / cnt, 'acc_' + key: 100 * acc_m / cnt, 'time_' + key: time.time() - t0})
        out['total_acc'] = (out['acc_public'] + out['acc_private']) / 2
        for k, v in test_signature(self.model).items():
            out['s_' + k] = v
        return out

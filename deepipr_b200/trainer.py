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
    """One optimisation step, V1 (one forward) or V2/V3 (public + private forward, one backward)."""

    def __init__(self, model, optimizer, private, buckets: GradBuckets = None, autocast=True):
        self.model, self.optimizer, self.private, self.buckets = model, optimizer, private, buckets
        self.autocast = autocast
        self._losses = sign_loss_modules(model)

    def forward_backward(self, data, target):
        """Returns (loss, sign_loss, [logits per pass]) as device tensors — no host sync in here."""
        self.optimizer.zero_grad()
        for m in self._losses:
            m.reset()
        preds = []
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.autocast):
            if self.private:
                loss = torch.zeros((), device=data.device)
                for ind in range(2):                               # "backprop to two graph at once"
                    pred = self.model(data, ind=ind)
                    loss = loss + F.cross_entropy(pred.float(), target)
                    preds.append(pred)
            else:
                pred = self.model(data)
                loss = F.cross_entropy(pred.float(), target)
                preds.append(pred)
        sign_loss = torch.zeros((), device=data.device)
        for m in self._losses:
            sign_loss = sign_loss + m.loss
        (loss + sign_loss).backward()
        if self.buckets is not None:
            self.buckets.finish()
        return loss, sign_loss, preds

    def step(self, data, target):
        loss, sign_loss, preds = self.forward_backward(data, target)
        self.optimizer.step()
        return loss, sign_loss, preds


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

    def __init__(self, model, optimizer, scheduler, device, buckets=None, autocast=True, verbose=False):
        self.model, self.optimizer, self.scheduler, self.device = model, optimizer, scheduler, device
        self.runner = StepRunner(model, optimizer, self.private, buckets, autocast)
        self.verbose = verbose

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
        self.model.train()
        sign_m = loss_m = acc_m = 0.0
        wm_iter = iter(wm_dataloader) if wm_dataloader is not None else None
        t0 = time.time()
        n = 0
        for data, target in dataloader:
            data = data.to(self.device, non_blocking=True)
            target = target.to(self.device, non_blocking=True)
            if wm_iter is not None:
                data, target, wm_iter = _cat_trigger(data, target, wm_iter, wm_dataloader, self.device)
            loss, sign_loss, preds = self.runner.step(data, target)
            sign_m += sign_loss.item()
            loss_m += loss.item()
            acc_m += accuracy(preds[0], target)[0].item()
            n += 1
        n = max(n, 1)
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
        self.model.train()
        loss_m = sign_m = pub_m = priv_m = 0.0
        wm_iter = iter(wm_dataloader) if wm_dataloader is not None else None
        t0 = time.time()
        n = 0
        for data, target in dataloader:
            data = data.to(self.device, non_blocking=True)
            target = target.to(self.device, non_blocking=True)
            if wm_iter is not None:
                data, target, wm_iter = _cat_trigger(data, target, wm_iter, wm_dataloader, self.device)
            loss, sign_loss, preds = self.runner.step(data, target)
            pub_m += accuracy(preds[0], target)[0].item()
            priv_m += accuracy(preds[1], target)[0].item()
            sign_m += sign_loss.item()
            loss_m += loss.item()
            n += 1
        n = max(n, 1)
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

"""World-size-2 data-parallel plumbing on CPU (gloo): flat gradient buckets must reproduce the gradient of the
concatenated global batch, with and without hook-driven overlap, and state broadcast must align the replicas.
(The passport kernels themselves are CUDA-only; this covers the host-side N>1 logic.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepipr_b200.parallel import FlatParams, GradBuckets, broadcast_state


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(12, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(),
                               torch.nn.Linear(32, 5))


def _worker(rank, world, port, overlap, bucket_bytes, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model(seed=100 + rank)            # replicas start different on purpose
        broadcast_state(model, src=0)
        flat = FlatParams(model.parameters())
        buckets = GradBuckets(flat, bucket_bytes=bucket_bytes, overlap=overlap)
        opt = torch.optim.SGD(flat.params, lr=0.1)
        g = torch.Generator().manual_seed(7)
        X = torch.randn(8, 12, generator=g)
        Y = torch.randint(0, 5, (8,), generator=g)
        xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]
        for step in range(2):
            opt.zero_grad()                        # set_to_none=True: the bucket views must be restored
            torch.nn.functional.cross_entropy(model(xs), ys).backward()
            buckets.finish()
            grads = [p.grad.clone() for p in flat.params]
            opt.step()
        out[rank] = dict(grads=grads, params=[p.detach().clone() for p in flat.params],
                         nbuckets=len(buckets.buckets),
                         views=all(p.grad.data_ptr() == flat.grad_view(i).data_ptr() for i, p in enumerate(flat.params)))
    finally:
        dist.destroy_process_group()


def _run(overlap, bucket_bytes):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, overlap, bucket_bytes, out), nprocs=2, join=True)
    # single-process reference on the global batch (mean over 8 == mean of the two per-rank means)
    model = _model(seed=100)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(7)
    X = torch.randn(8, 12, generator=g)
    Y = torch.randint(0, 5, (8,), generator=g)
    for step in range(2):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(X), Y).backward()
        ref_grads = [p.grad.clone() for p in model.parameters()]
        opt.step()
    for rank in (0, 1):
        r = out[rank]
        assert r["views"]
        for a, b in zip(r["grads"], ref_grads):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
        for a, b in zip(r["params"], model.parameters()):
            assert torch.allclose(a, b.detach(), rtol=1e-5, atol=1e-6)
    return out[0]["nbuckets"]


def test_bucketed_allreduce_overlap_many_buckets():
    assert _run(overlap=True, bucket_bytes=1024) > 2


def test_bucketed_allreduce_no_overlap_single_bucket():
    assert _run(overlap=False, bucket_bytes=1 << 30) == 1


def test_flat_params_keeps_values_and_views():
    model = _model(seed=1)
    before = [p.detach().clone() for p in model.parameters()]
    flat = FlatParams(model.parameters())
    for p, b in zip(flat.params, before):
        assert torch.equal(p.detach(), b)
    flat.flat.mul_(2)
    for p, b in zip(flat.params, before):
        assert torch.equal(p.detach(), b * 2)          # parameters are views of the flat buffer
    model(torch.randn(3, 12)).sum().backward()
    assert flat.flat_grad.abs().sum() > 0              # gradients landed in the flat gradient buffer
    flat.zero_grad()
    assert flat.flat_grad.abs().sum() == 0 and all(p.grad is not None for p in flat.params)


# ---- directly accumulated gradients (what functional._ConvBlockFn does on the GPU): the operator adds its weight
# gradient into FlatParams.flat_grad itself, returns None to autograd and reports readiness through
# FlatParams.direct_done().  Autograd still runs the parameter's AccumulateGrad node (with an undefined gradient) and
# fires its post-accumulate hook; counting that hook as a second readiness signal launched buckets early.
class _DirectLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, flat, i):
        ctx.save_for_backward(x, w)
        ctx.flat, ctx.i = flat, i
        flat.direct_begin(i)
        return x @ w.t()

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        ctx.flat.grad_view(ctx.i).add_(gy.t() @ x)
        ctx.flat.direct_done(ctx.i)
        return gy @ w, None, None, None


def _two_pass_loss(ws, x, y, lin):
    loss = 0
    for scale in (1.0, 0.5):                       # two passes over the same weights, one backward (V2 step shape)
        h = torch.relu(lin(0, x * scale, ws[0]))
        h = torch.relu(lin(1, h, ws[1]))
        h = torch.relu(lin(2, h, ws[2]))
        loss = loss + torch.nn.functional.cross_entropy(torch.nn.functional.linear(h, ws[3]), y)
    return loss


def _weights():
    g = torch.Generator().manual_seed(3)
    return [torch.nn.Parameter(torch.randn(s, generator=g) * 0.3) for s in ((32, 12), (32, 32), (32, 32), (5, 32))]


def _direct_worker(rank, world, port, bucket_bytes, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ws = _weights()
        flat = FlatParams(ws)
        buckets = GradBuckets(flat, bucket_bytes=bucket_bytes, overlap=True)
        g = torch.Generator().manual_seed(7)
        X, Y = torch.randn(8, 12, generator=g), torch.randint(0, 5, (8,), generator=g)
        xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]
        grads = []
        for step in range(2):
            flat.zero_grad()
            _two_pass_loss(ws, xs, ys, lambda i, x, w: _DirectLinear.apply(x, w, flat, i)).backward()
            buckets.finish()
            grads.append(flat.flat_grad.clone())
        out[rank] = dict(grads=grads, nbuckets=len(buckets.buckets), offsets=list(flat.offsets))
    finally:
        dist.destroy_process_group()


def _run_direct(bucket_bytes, want_buckets):
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_direct_worker, args=(2, port, bucket_bytes, out), nprocs=2, join=True)
    ws = _weights()
    g = torch.Generator().manual_seed(7)
    X, Y = torch.randn(8, 12, generator=g), torch.randint(0, 5, (8,), generator=g)
    # mean over the global batch == mean of the per-rank means (equal shard sizes)
    loss = 0.5 * (_two_pass_loss(ws, X[:4], Y[:4], lambda i, x, w: torch.nn.functional.linear(x, w)) +
                  _two_pass_loss(ws, X[4:], Y[4:], lambda i, x, w: torch.nn.functional.linear(x, w)))
    loss.backward()
    assert out[0]["nbuckets"] == want_buckets
    for rank in (0, 1):
        for step in range(2):
            for w, o in zip(ws, out[rank]["offsets"]):
                got = out[rank]["grads"][step][o:o + w.numel()].view_as(w)
                assert torch.allclose(got, w.grad, rtol=1e-5, atol=1e-6)


def test_directly_accumulated_gradients_shared_bucket():
    # all four parameters in ONE bucket: with the hook double-counted it was launched after two of the three
    # direct parameters, before the first layer's gradient existed
    _run_direct(bucket_bytes=1 << 16, want_buckets=1)


def test_directly_accumulated_gradients_many_buckets():
    _run_direct(bucket_bytes=6000, want_buckets=2)


# ---- gradient accumulation / repeated backward: a second backward before finish() must not add
# un-reduced gradients on top of buckets that are already divided and in flight
def _accum_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model(seed=100)
        flat = FlatParams(model.parameters())
        buckets = GradBuckets(flat, bucket_bytes=1024, overlap=True)
        g = torch.Generator().manual_seed(7)
        X = torch.randn(16, 12, generator=g)
        Y = torch.randint(0, 5, (16,), generator=g)
        xs, ys = X[rank * 8:(rank + 1) * 8], Y[rank * 8:(rank + 1) * 8]
        ce = torch.nn.functional.cross_entropy
        # (1) two backward passes without no_sync(): refused, not silently wrong
        flat.zero_grad()
        ce(model(xs[:4]), ys[:4]).backward()
        try:
            ce(model(xs[4:]), ys[4:]).backward()
            refused = False
        except RuntimeError as e:
            refused = "no_sync" in str(e)
        buckets.finish()
        # (2) the supported way: all but the last micro-batch inside no_sync()
        flat.zero_grad()
        with buckets.no_sync():
            (0.5 * ce(model(xs[:4]), ys[:4])).backward()
        (0.5 * ce(model(xs[4:]), ys[4:])).backward()
        buckets.finish()
        accum = flat.flat_grad.clone()
        # (3) a parameter without a gradient after zero_grad(set_to_none=True) contributes zeros, not stale values
        opt = torch.optim.SGD(flat.params, lr=0.1)
        flat.flat_grad.fill_(123.0)
        opt.zero_grad()                              # set_to_none=True
        head = list(model.children())[-1]
        ce(head(torch.randn(4, 32, generator=g)), ys[:4]).backward()     # only the last layer gets a gradient
        buckets.finish()
        first_w = flat.grad_view(0).clone()
        out[rank] = dict(refused=refused, accum=accum, first_w=first_w, offsets=list(flat.offsets))
    finally:
        dist.destroy_process_group()


def test_gradient_accumulation_no_sync_and_second_backward_is_refused():
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_accum_worker, args=(2, port, out), nprocs=2, join=True)
    model = _model(seed=100)
    g = torch.Generator().manual_seed(7)
    X = torch.randn(16, 12, generator=g)
    Y = torch.randint(0, 5, (16,), generator=g)
    torch.nn.functional.cross_entropy(model(X), Y).backward()      # mean over the global batch of 16
    for rank in (0, 1):
        r = out[rank]
        assert r["refused"], "a second backward before finish() must raise"
        for p, o in zip(model.parameters(), r["offsets"]):
            assert torch.allclose(r["accum"][o:o + p.numel()].view_as(p), p.grad, rtol=1e-5, atol=1e-6)
        assert r["first_w"].abs().sum() == 0, "stale gradient of a parameter that got no gradient was reduced"


def test_flat_sgd_refuses_a_second_param_group_and_round_trips_its_state():
    from deepipr_b200.parallel import FlatSGD
    import pytest
    ps = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    flat = FlatParams(ps)
    opt = FlatSGD(flat, lr=0.1, momentum=0.9, weight_decay=1e-4)
    with pytest.raises(ValueError, match="param group"):
        opt.add_param_group({"params": [torch.nn.Parameter(torch.randn(2))]})
    opt._buf.normal_()
    opt._steps = 7
    sd = opt.state_dict()
    opt2 = FlatSGD(FlatParams([torch.nn.Parameter(p.detach().clone()) for p in ps]), lr=0.5)
    opt2.load_state_dict(sd)
    assert opt2._steps == 7 and torch.equal(opt2._buf, opt._buf) and opt2.param_groups[0]["lr"] == 0.1

#!/usr/bin/env python
"""Benchmark of the passport-layer training path (contract: the task brief / DESIGN.md "Measurement").

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 20 --warmup 5       # the UNMODIFIED reference trainer on the host cores
    python bench.py --config v1_imagenet | v2_cifar100 | v1_alexnet | v2_cifar10     # the other BASELINE configs

Default workload (BASELINE.json metric "images/sec ResNet18-passport CIFAR10 train", configs[3]): one
TrainerPrivate step — public + private forward, one backward, SGD — of ResNet-18 with passport layers in layer4
(passport_configs/resnet18_passport.json) on synthetic CIFAR-10-shaped tensors, V3: every step appends 2 trigger-set
images (experiments/trainer_private.py:135-146), bf16 activations, per-GPU batch 1024, weak scaling.
One JSON line on stdout (rank 0).
"""
import argparse
import contextlib
import ctypes as C
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec ResNet18-passport CIFAR10 train"
UNIT = "images/s"

#: the BASELINE.json configurations (2..5) + the round-1 workload; gflop = algorithmic GFLOP per image-step
CONFIGS = {
    "v3_cifar10_trigger": dict(net="resnet18", scheme="private", classes=10, img=32, batch=1024, trigger=True,
                               gflop=6.665, dtype="bf16",
                               workload="ResNet18 V3 private-passport + trigger set (layer4: 5 passport convs), "
                                        "CIFAR10-shaped TrainerPrivate step (BASELINE configs[3])"),
    "v2_cifar10": dict(net="resnet18", scheme="private", classes=10, img=32, batch=1024, trigger=False, gflop=6.665,
                       dtype="bf16", workload="ResNet18 V2 private-passport, CIFAR10-shaped TrainerPrivate step "
                                              "(round-1 workload)"),
    "v2_cifar100": dict(net="resnet18", scheme="private", classes=100, img=32, batch=1024, trigger=False, gflop=6.665,
                        dtype="bf16", workload="ResNet18 V2 private-passport CIFAR100-shaped TrainerPrivate step "
                                               "(BASELINE configs[2])"),
    "v1_imagenet": dict(net="resnet18", scheme="v1", classes=1000, img=224, batch=256, trigger=False, gflop=10.884,
                        dtype="bf16", workload="ResNet18 V1 passport ImageNet-1k-shaped Trainer step, 7x7/s2 stem + "
                                               "max-pool, lr_configs/imagenet.json schedule (BASELINE configs[4])"),
    # fp32 tensors end to end, contractions on tcgen05 kind::tf32 (PP_DTYPE_TF32) — what train_v1.py's fp32 run gets
    # from cuDNN on a GPU; the TF32 tensor peak is half the bf16 one
    "v1_alexnet": dict(net="alexnet", scheme="v1", classes=10, img=32, batch=1024, trigger=False, gflop=1.323,
                       dtype="tf32", workload="AlexNet V1 passport (features 4/5/6) CIFAR10-shaped Trainer step, fp32 "
                                              "activations / TF32 tensor cores (BASELINE configs[1])"),
    "v1_alexnet_bf16": dict(net="alexnet", scheme="v1", classes=10, img=32, batch=1024, trigger=False, gflop=1.323,
                            dtype="bf16", workload="AlexNet V1 passport (features 4/5/6) CIFAR10-shaped Trainer step "
                                                   "under autocast(bf16) (not a BASELINE config: precision below the "
                                                   "reference's fp32)"),
}
DEFAULT_CONFIG = "v3_cifar10_trigger"
CPU_SAMPLE_BATCH = 256             # images of the per-step batch the CPU arms process (bounded sample, see below)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops_sustained", 1393.0), tflops_burst=d.get("bf16_tflops", 1665.0),
                    hbm=d.get("hbm_gbs", 6540.0), src="measured")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, src="fallback")


def tensor_peak(pk, dtype):
    """Tensor-core peak (TFLOP/s) for a config's arithmetic: the measured sustained bf16 figure; kind::tf32 runs at half
    the bf16 rate on this part (8 instead of 16 k per instruction at the same issue cost)."""
    return pk["tflops"] * (0.5 if dtype == "tf32" else 1.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        """Summary of the samples taken inside [t_begin, t_end] (the timed region); if the region was shorter than the
        sampling period, of the samples taken under load since start() (warm-up steps run the same kernels)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def summarise(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[1])); mx.append(float(r[2]))
                    for name, v in zip(names, r[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
                except Exception:
                    pass
            return sm, mx, reasons

        window = [row for row in self.rows if t_begin is None or (t_begin <= row[0] <= (t_end or 1e18) + 0.1)]
        scope = "timed region"
        sm, mx, reasons = summarise(window)
        if not sm:                                  # region shorter than one sampling period
            loaded = []
            for row in self.rows:
                try:
                    if float(row[1][3]) > 400.0:      # power draw: the device was running the step
                        loaded.append(row)
                except Exception:
                    pass
            sm, mx, reasons = summarise(loaded or self.rows)
            scope = "warm-up + timed region (samples under load)"
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


def _seed(seed):
    import random
    import numpy as np
    import torch
    torch.manual_seed(seed); random.seed(seed); np.random.seed(seed)


def build_model(num_classes=None, seed=0, blocks=None, config=None):
    """The product network of a config (deepipr_b200.nets wiring), passports fixed up front (key_type='random'
    semantics: U(-1,1), passportconv2d.py:198-207) so the lazily-created-key branch is not in the timed region."""
    import numpy as np
    import torch
    from deepipr_b200 import nets
    cfg = CONFIGS[config or "v2_cifar10"]
    classes = num_classes if num_classes is not None else cfg["classes"]
    _seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        if cfg["net"] == "alexnet":
            pk = nets.passport_kwargs_from_config(nets.alexnet_passport_config(), "bn", "random", 0.1)
            model = nets.AlexNetCifar(cfg["scheme"], 3, classes, pk)
        else:
            pk = nets.passport_kwargs_from_config(nets.resnet18_passport_config(), "bn", "random", 0.1)
            model = nets.ResNet18(cfg["scheme"], classes, pk)
    imagenet = cfg["img"] == 224
    for m in model.modules():
        if getattr(m, "KIND", None) in ("private", "v1"):
            c = m.conv.in_channels
            if cfg["net"] == "alexnet":
                h = 8
            else:
                wide = cfg["img"] // (16 if imagenet else 4)              # input of layer4.0.convbnrelu_1 / shortcut
                h = wide if m.conv.stride[0] == 2 else wide // 2
            m.set_key(torch.tensor(np.random.uniform(-1, 1, (1, c, h, h)), dtype=torch.float32),
                      torch.tensor(np.random.uniform(-1, 1, (1, c, h, h)), dtype=torch.float32))
    return model


def synthetic_batches(cfg, batch, n, device=None, seed=1234):
    import torch
    g = torch.Generator(device=device or "cpu").manual_seed(seed)
    return [(torch.randn(batch, 3, cfg["img"], cfg["img"], device=device, generator=g),
             torch.randint(0, cfg["classes"], (batch,), device=device, generator=g)) for _ in range(n)]


def trigger_batches(n=50, device=None, seed=4321):
    """prepare_wm: 100 trigger images, batch 2, CIFAR labels (dataset.py:168-193)."""
    import torch
    g = torch.Generator(device=device or "cpu").manual_seed(seed)
    return [(torch.randn(2, 3, 32, 32, device=device, generator=g), torch.randint(0, 10, (2,), device=device,
                                                                                  generator=g)) for _ in range(n)]


# ------------------------------------------------------------------------------------------------------------------
# baseline arms: the UNMODIFIED reference (oracle/_ref bundle or checkout), never the product path
# ------------------------------------------------------------------------------------------------------------------
def reference_trainer(cfg, device, channels_last=False):
    """(model, trainer, kind): the reference's own model class, SGD and Trainer / TrainerPrivate
    (experiments/classification_private.py:48-62), built from its passport_configs/*.json."""
    import torch
    from oracle import ref_bundle
    mods = ref_bundle.import_reference(patched=False)
    ref = ref_bundle.locate()
    name = "alexnet_passport.json" if cfg["net"] == "alexnet" else "resnet18_passport.json"
    pcfg = json.load(open(os.path.join(ref, "passport_configs", name)))
    pk = mods["experiments.utils"].construct_passport_kwargs_from_dict(
        {"passport_config": pcfg, "norm_type": "bn", "key_type": "random", "sl_ratio": 0.1})
    _seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        if cfg["net"] == "alexnet":
            model = mods["models.alexnet_passport"].AlexNetPassport(3, cfg["classes"], pk)
        elif cfg["scheme"] == "private":
            model = mods["models.resnet_passport_private"].ResNet18Private(num_classes=cfg["classes"],
                                                                           passport_kwargs=pk)
        else:
            model = mods["models.resnet_passport"].ResNet18Passport(num_classes=cfg["classes"], passport_kwargs=pk)
    model = model.to(device)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    # The reference's trainers wrap the model in nn.DataParallel whenever torch.cuda.device_count() > 1
    # (experiments/trainer.py:92-93, trainer_private.py:110-111) — also for a CPU model, which then fails.  This arm is
    # one process on one device (the CPU, or one GPU): construct the trainer the way CUDA_VISIBLE_DEVICES=<one device>
    # would make it see the machine.  Nothing in the reference is modified.
    real_count = torch.cuda.device_count
    torch.cuda.device_count = lambda: min(1, real_count())
    try:
        if cfg["scheme"] == "private":
            trainer = mods["experiments.trainer_private"].TrainerPrivate(model, opt, None, device)
        else:
            trainer = mods["experiments.trainer"].Trainer(model, opt, None, device)
    finally:
        torch.cuda.device_count = real_count
    return model, trainer


def reference_run(cfg, steps, warmup, device, batch, autocast=False, channels_last=False):
    """images/s of the reference's own trainer.train() over `steps` minibatches (after `warmup`), on `device`."""
    import torch
    model, trainer = reference_trainer(cfg, device, channels_last)
    data = synthetic_batches(cfg, batch, 4, device=None)
    if str(device) != "cpu":
        data = [(x.to(device), t.to(device)) for x, t in data]
        if channels_last:
            data = [(x.contiguous(memory_format=torch.channels_last), t) for x, t in data]
    wm = trigger_batches(8) if cfg["trigger"] else None
    if wm is not None and str(device) != "cpu":
        wm = [(x.to(device), t.to(device)) for x, t in wm]
    ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if autocast else contextlib.nullcontext

    def epoch(n):
        loader = [data[i % 4] for i in range(n)]
        with contextlib.redirect_stdout(io.StringIO()), ctx():
            return trainer.train(0, loader, wm)

    epoch(max(1, warmup))
    if str(device) != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    epoch(steps)
    if str(device) != "cpu":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    per_step = batch + (2 if cfg["trigger"] else 0)
    return steps * per_step / dt, dt / steps


def cpu_reference_arm(cfg, steps, warmup):
    import torch
    from oracle import ref_bundle
    torch.set_num_threads(os.cpu_count() or 1)
    sample = min(CPU_SAMPLE_BATCH, cfg["batch"]) if cfg["img"] == 32 else 16
    if ref_bundle.available():
        value, sec = reference_run(cfg, steps, warmup, torch.device("cpu"), sample)
        kind, what = "reference", "the unmodified reference (oracle/_ref bundle): its model class, torch.optim.SGD and " \
                                  "Trainer/TrainerPrivate.train"
    else:                                    # no bundle on this machine: the oracle port of the same loop
        from oracle import passport_oracle as po
        model = po.mirror(build_model(config=_config_name(cfg)), round_bf16=False).train()
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
        data = synthetic_batches(cfg, sample, 4)
        for i in range(max(1, warmup)):
            po.train_step(model, opt, *data[i % 4], private=cfg["scheme"] == "private")
        t0 = time.perf_counter()
        for i in range(steps):
            po.train_step(model, opt, *data[i % 4], private=cfg["scheme"] == "private")
        sec = (time.perf_counter() - t0) / steps
        value, kind, what = sample / sec, "port", "oracle port of experiments/trainer_private.py:148-177"
    return dict(value=value, unit=UNIT, cores=torch.get_num_threads(), kind=kind,
                sample=f"{steps} steps (+{max(1, warmup)} warm-up) of {what}; each step a bounded sample of the "
                       f"workload's per-step batch: {sample} of {cfg['batch']} images"
                       f"{' + 2 trigger images' if cfg['trigger'] else ''}, fp32, device=cpu"), sec


def _config_name(cfg):
    return next(k for k, v in CONFIGS.items() if v is cfg)


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG, choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=5)
    ap.add_argument("--graph", action="store_true", help="(default) value / e2e legs replay the step as one CUDA graph")
    ap.add_argument("--no-graph", action="store_true", help="eager launches (~540 per step) instead of the graph replay")
    ap.add_argument("--legs", default="value,e2e,roofline,eager,dropin,small_batch,configs,shared",
                    help="comma list: value (always), e2e, roofline, eager (the reference on this GPU), dropin (the "
                         "reference's trainer on the patched blocks), small_batch (batch 64 / 256, eager vs CUDA "
                         "graph), configs (short lines of the other BASELINE configs), shared (trunk CSE)")
    args = ap.parse_args()

    # stdout carries exactly ONE JSON line.  Native libraries write to file descriptor 1 behind Python's back (NCCL
    # prints its version banner there): keep the real stdout aside for the JSON line and point fd 1 at stderr.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = CONFIGS[args.config]
    B = args.batch or cfg["batch"]
    per_step = B + (2 if cfg["trigger"] else 0)           # images a rank processes per step

    config = {"workload": cfg["workload"], "name": args.config, "per_gpu_batch": B,
              "trigger_images_per_step_per_gpu": 2 if cfg["trigger"] else 0,
              "global_batch": per_step * max(world, 1), "parallelism": f"dp{max(world, 1)}",
              "passport_layers": "features 4/5/6" if cfg["net"] == "alexnet" else "layer4 (5 convs)", "norm": "bn",
              "classes": cfg["classes"], "image": cfg["img"],
              "optimizer": "SGD(0.01, momentum 0.9, wd 1e-4), fused flat step",
              "l2": "per-step working set (GBs of activations) >> 126 MB L2; 4 rotating input batches; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        cb, sec_per_step = cpu_reference_arm(cfg, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from deepipr_b200 import _lib as L
    from deepipr_b200.parallel import FlatParams, FlatSGD, GradBuckets, broadcast_state
    from deepipr_b200.trainer import GraphedStepRunner, StepRunner, Trainer, TrainerPrivate, test_signature

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own logging goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()
    legs = set(args.legs.split(","))
    pk = peaks()
    private = cfg["scheme"] == "private"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    def make_runner(config_name, batch, graph=False, ddp=True):
        """model + flat SGD + (graphed) step runner for one config; inputs resident in HBM."""
        c = CONFIGS[config_name]
        from deepipr_b200 import layers as _layers
        _layers.set_precision("tf32" if c["dtype"] == "tf32" else "bf16")
        model = build_model(config=config_name).to(dev).train()
        if world > 1 and ddp:
            broadcast_state(model)
        flat = FlatParams(model.parameters())
        opt = FlatSGD(flat, lr=0.01, momentum=0.9, weight_decay=1e-4)
        buckets = GradBuckets(flat) if (world > 1 and ddp) else None
        runner = StepRunner(model, opt, private=c["scheme"] == "private", buckets=buckets,
                            autocast=c["dtype"] == "bf16")
        data = synthetic_batches(c, batch, 4, device=dev, seed=1234 + rank)
        if c["trigger"]:
            wm = trigger_batches(4, device=dev, seed=4321 + rank)
            data = [(torch.cat([x, wm[i][0]]), torch.cat([t, wm[i][1]])) for i, (x, t) in enumerate(data)]
        step = runner.step
        if graph:
            graphed = GraphedStepRunner(runner, *data[0])
            step = graphed.step
            runner.graphed = graphed
        return model, opt, runner, data, step

    def timed(step, data, steps, warmup):
        for i in range(warmup):
            step(*data[i % len(data)])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(*data[i % len(data)])
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---------------- leg 1 (`value`): inputs resident in HBM
    # The step (zero_grad, forward(s), losses, backward, bucketed NCCL all-reduces when N > 1, fused SGD) is captured
    # once and replayed: trainer.GraphedStepRunner.  --no-graph runs the same step eagerly.
    use_graph = not args.no_graph
    model, opt, runner, dev_batches, step = make_runner(args.config, B, graph=use_graph)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                 # nvidia-smi needs a few hundred ms to deliver its first sample
    for i in range(args.warmup):
        step(*dev_batches[i % 4])
    barrier()
    lib.pp_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    e0.record()
    for i in range(args.steps):
        step(*dev_batches[i % 4])
    e1.record()
    barrier()
    t_end = time.time()
    launches = int(lib.pp_launch_count(0))
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    value = args.steps * per_step * world / (ms_total * 1e-3)
    if use_graph:       # a replay re-launches the kernels counted while the step was captured
        launches = runner.graphed.launches_per_replay * args.steps

    # ---------------- leg 2 (`e2e`): the public trainer API on HOST buffers.  deepipr_b200.trainer.TrainerPrivate.train
    # (the call a user makes, same signature as experiments/trainer_private.py:118) over a loader of pinned host
    # batches + the trigger-set loader: every step copies its inputs host->device (prefetched on a side stream, as
    # DataLoader(pin_memory=True) + .to(non_blocking=True) allows) and reads the step's metrics back (16 bytes).
    ms_e2e, e2e_value, last = None, None, None
    trainer_cls = TrainerPrivate if private else Trainer
    host = synthetic_batches(cfg, B, 4, device=None, seed=1234 + rank)
    host = [(x.pin_memory(), t.pin_memory()) for x, t in host]
    wm_host = [(x.pin_memory(), t.pin_memory()) for x, t in trigger_batches(8, seed=4321 + rank)] \
        if cfg["trigger"] else None
    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 8 + \
        ((wm_host[0][0].numel() * 4 + wm_host[0][1].numel() * 8) if wm_host else 0)
    if "e2e" in legs:
        trainer = trainer_cls(model, opt, None, dev, buckets=runner.buckets, autocast=cfg["dtype"] == "bf16",
                              use_graph=use_graph)
        trainer.log_every = 1
        seen = []
        trainer.on_log = lambda n, vals: seen.append(vals)
        trainer.train(0, [host[i % 4] for i in range(max(3, args.warmup // 2))], wm_host)
        barrier()
        seen.clear()
        e0.record()
        trainer.train(1, [host[i % 4] for i in range(args.steps)], wm_host)
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        e2e_value = args.steps * per_step * world / (ms_e2e * 1e-3)
        n = len(seen)
        last = {"loss": seen[-1][0] / args.steps, "sign_loss_sum": seen[-1][1], "acc_pass0": seen[-1][2] / args.steps,
                "acc_pass1": seen[-1][3] / args.steps, "metric_reads": n}
    d2h = 4 * 4

    # ---------------- leg 3 (`roofline`): per-kernel CUDA-event timing on the kernels' own stream
    roof = roof_w = roof_hbm = roof_fused = None
    if "roofline" in legs:                   # (eager steps: the per-kernel event instrumentation wraps real launches)
        # one stream for these steps: a kernel that shares the SMs with another stream's kernel (the side-stream weight
        # gradients) would be charged the other one's time by the per-launch events
        from deepipr_b200 import functional as _F
        overlap_was, _F.OVERLAP_WGRAD = _F.OVERLAP_WGRAD, False
        if rank == 0:
            lib.pp_profile_enable(1)
        for i in range(3):
            runner.step(*dev_batches[i % 4])
        barrier()
        lib.pp_profile_enable(0)
        _F.OVERLAP_WGRAD = overlap_was
    if rank == 0 and "roofline" in legs:
        def read(kind, c=0, nout=0, taps=0):
            ms, fl, n = C.c_double(0), C.c_double(0), C.c_int(0)
            lib.pp_profile_read(kind, c, nout, taps, C.byref(ms), C.byref(fl), C.byref(n))
            return ms.value, fl.value, n.value

        traffic_file = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
        ncu_traffic = json.load(open(traffic_file)) if os.path.exists(traffic_file) else {}
        step_ms = ms_total / args.steps
        (ms0, fl0, n0), (ms1, fl1, n1) = read(0), read(1)
        if ms0 > 0:
            ach = fl0 / (ms0 * 1e-3) / 1e12
            roof = {"kernel": "pxn_kernel + tapgemm_kernel (tcgen05 implicit-GEMM conv, fprop + dgrad launches)",
                    "bound": "tensor", "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s",
                    "frac": ach / pk["tflops"], "peak_source": pk["src"] + " bf16 sustained (kernel timed inside a step)",
                    "traffic": ncu_traffic.get("tapgemm_total_bytes_per_launch"),
                    "traffic_source": ncu_traffic.get("source"),
                    "launches_per_step": n0 // 3, "avg_launch_us": ms0 * 1e3 / max(n0, 1),
                    "share_of_step": (ms0 / 3) / step_ms}
        msp, flp, npl = read(0, 512, 512, 9)      # the 3x3 512->512 passport convs of layer4 (fprop + dgrad)
        if msp > 0 and roof is not None:
            ach = flp / (msp * 1e-3) / 1e12
            roof["passport_layer"] = {"geometry": "layer4 3x3 512->512, fprop+dgrad launches", "achieved": ach,
                                      "frac": ach / pk["tflops"], "launches_per_step": npl // 3,
                                      "avg_launch_us": msp * 1e3 / max(npl, 1),
                                      "algorithmic_bytes": 2.0 * (per_step * 16 * 512 + 512 * 4608),
                                      "traffic": ncu_traffic.get("passport_layer_bytes_per_launch")}
        msf, flf, nf = read(5)
        if msf > 0:
            ach = flf / (msf * 1e-3) / 1e12
            roof_fused = {"kernel": "passport_fused_kernel (conv + batch statistics + grid barrier + gamma/beta affine + "
                                    "ReLU from TMEM, one launch per passport block forward)", "bound": "tensor",
                          "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                          "launches_per_step": nf // 3, "avg_launch_us": msf * 1e3 / max(nf, 1),
                          "share_of_step": (msf / 3) / step_ms,
                          "traffic": ncu_traffic.get("passport_fused_bytes_per_launch")}
        hbm_parts, tot_ms, tot_b = {}, 0.0, 0.0
        for kind, name in ((2, "affine_apply (z->y)"), (3, "bwd reduce (dy,z)"), (4, "bwd dz (dy,z->dz)")):
            msk, byk, nk = read(kind)
            if msk > 0:
                hbm_parts[name] = {"achieved": byk / (msk * 1e-3) / 1e9, "launches_per_step": nk // 3,
                                   "avg_launch_us": msk * 1e3 / max(nk, 1), "share_of_step": (msk / 3) / step_ms}
                tot_ms += msk
                tot_b += byk
        if tot_ms > 0:
            ach = tot_b / (tot_ms * 1e-3) / 1e9
            roof_hbm = {"kernel": "affine_apply + column_reduce<1> + bwd_dz (norm/affine/ReLU passes)", "bound": "hbm",
                        "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                        "peak_source": pk["src"] + " copy bandwidth", "traffic": ncu_traffic.get("hbm_passes"),
                        "traffic_source": ncu_traffic.get("source"),
                        "share_of_step": (tot_ms / 3) / step_ms, "per_kernel": hbm_parts,
                        "note": "algorithmic bytes as launched (fp32 z); small layers (layer3/4) re-read z/dy from the "
                                "126 MB L2, so a per-kernel figure can exceed the HBM peak"}
        if ms1 > 0:
            ach = fl1 / (ms1 * 1e-3) / 1e12
            roof_w = {"kernel": "wgrad_kernel / wgrad_om_kernel (tcgen05, MN-major)", "bound": "tensor", "achieved": ach,
                      "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                      "launches_per_step": n1 // 3, "avg_launch_us": ms1 * 1e3 / max(n1, 1),
                      "share_of_step": (ms1 / 3) / step_ms}

    # ---------------- optional: public/private passes share the passport-free trunk (reported separately)
    shared = None
    if "shared" in legs and private and cfg["net"] == "resnet18":
        model.share_trunk = True
        ms_sh = timed(runner.step, dev_batches, max(5, args.steps // 2), 3)
        n_sh = max(5, args.steps // 2)
        shared = {"value": n_sh * per_step * world / (ms_sh * 1e-3), "unit": UNIT, "ms_per_step": ms_sh / n_sh,
                  "what": "public+private passes share the stem..layer3 forward/backward (common-subexpression "
                          "elimination across the two model calls of trainer_private.py:159-161); opt-in, not `value`"}
        model.share_trunk = False

    sig = test_signature(model) if rank == 0 else {}
    model.train()
    if getattr(runner, "graphed", None) is not None:
        runner.graphed.release()
    del model, opt, runner, dev_batches, step
    torch.cuda.empty_cache()

    # ---------------- the other BASELINE configs, short lines (N=1 only; the scaling run repeats the default config)
    configs_out = None
    if "configs" in legs and world == 1:
        configs_out = {}
        for name in ("v2_cifar100", "v1_imagenet", "v1_alexnet"):
            if name == args.config:
                continue
            c = CONFIGS[name]
            try:
                m2, o2, r2, d2, s2 = make_runner(name, c["batch"], graph=use_graph, ddp=False)
                n2 = 8
                ms2 = timed(s2, d2, n2, 3)
                v2 = n2 * c["batch"] / (ms2 * 1e-3)
                if getattr(r2, "graphed", None) is not None:
                    r2.graphed.release()
                configs_out[name] = {"value": v2, "unit": UNIT, "ms_per_step": ms2 / n2, "per_gpu_batch": c["batch"],
                                     "dtype": c["dtype"], "workload": c["workload"],
                                     "conv_roofline_frac_whole_step":
                                         v2 * c["gflop"] * 1e9 / (tensor_peak(pk, c["dtype"]) * 1e12)}
                del m2, o2, r2, d2, s2
            except Exception as e:          # a config must not take the headline line down with it
                configs_out[name] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
        from deepipr_b200 import layers as _layers
        _layers.set_precision("tf32" if cfg["dtype"] == "tf32" else "bf16")

    # ---------------- throughput at the reference's own batch sizes (64: train_v1.py:15, 256: training.sh:4):
    # ~650 launches per step make the eager step host-bound there; the CUDA-graph replay is the fix
    small = None
    if "small_batch" in legs and world == 1:
        small = {}
        for b in (64, 256):
            row = {}
            for mode in ("eager", "graph"):
                try:
                    m2, o2, r2, d2, s2 = make_runner(args.config, b, graph=(mode == "graph"), ddp=False)
                    n2 = 20
                    ms2 = timed(s2, d2, n2, 5)
                    row[mode] = {"value": n2 * (b + (2 if cfg["trigger"] else 0)) / (ms2 * 1e-3), "unit": UNIT,
                                 "ms_per_step": ms2 / n2}
                    del m2, o2, r2, d2, s2
                except Exception as e:
                    row[mode] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()
            small[f"batch_{b}"] = row

    # ---------------- the reference itself on this GPU (stock PyTorch eager + cuDNN): "the honest bar"
    eager = None
    multi_dev = torch.cuda.device_count() > 1     # the reference trainers wrap nn.DataParallel when they see >1 GPU
    if rank == 0 and world == 1 and "eager" in legs:
        from oracle import ref_bundle
        if multi_dev:
            eager = {"skipped": "more than one CUDA device visible: the reference would switch to nn.DataParallel"}
        elif ref_bundle.available():
            eager = {}
            torch.backends.cudnn.benchmark = True                      # train_v1.py:8 / train_v23.py:8
            for tag, kw in (("fp32_as_shipped", dict()),
                            ("autocast_bf16_channels_last", dict(autocast=True, channels_last=True))):
                try:
                    n2 = max(3, min(args.steps, 6))
                    v, sec = reference_run(cfg, n2, 2, dev, B, **kw)
                    eager[tag] = {"value": v, "unit": UNIT, "ms_per_step": sec * 1e3, "steps": n2}
                except Exception as e:
                    eager[tag] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()
            eager["what"] = ("the UNMODIFIED reference (oracle/_ref bundle): its ResNet18Private, torch.optim.SGD and "
                             "TrainerPrivate.train on cuda, same per-GPU batch, inputs resident; fp32_as_shipped = what "
                             "`python train_v23.py` does (cuDNN TF32 convs by torch default); the second row wraps the "
                             "same call in torch.autocast(bf16) with a channels_last model")
        else:
            eager = {"unavailable": "no reference bundle (oracle/_ref) on this machine"}

    # ---------------- drop-in route: the reference's OWN TrainerPrivate + model files on the patched blocks
    dropin = None
    if rank == 0 and world == 1 and "dropin" in legs and not multi_dev:
        from oracle import ref_bundle
        if ref_bundle.available():
            try:
                dropin = dropin_run(cfg, B, dev, max(3, min(args.steps, 6)))
            except Exception as e:
                dropin = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = cpu_reference_arm(cfg, args.cpu_steps, 1)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic", "config": config,
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": (ms_e2e / args.steps) if ms_e2e else None,
                        "api": "deepipr_b200.trainer.TrainerPrivate.train(epoch, loader of pinned host batches, "
                               "trigger loader)"},
                "gpu_launches": launches, "cuda_graph": use_graph,
                "roofline": roof, "roofline_passport_fused": roof_fused, "roofline_wgrad": roof_w,
                "roofline_hbm": roof_hbm,
                "conv_roofline_frac_whole_step":
                    (value / world) * cfg["gflop"] * 1e9 / (tensor_peak(pk, cfg["dtype"]) * 1e12),
                "cpu_baseline": cpu_baseline, "torch_eager_gpu": eager, "reference_trainer_on_patched_blocks": dropin,
                "small_batch": small, "configs": configs_out, "value_shared_trunk": shared,
                "sign_bit_accuracy": (sum(sig.values()) / len(sig)) if sig else None, "last_step": last}
        emit(line)
    # teardown: captured graphs hold references into the NCCL communicator, so they go first; and a communicator
    # teardown that does not return must not turn a finished measurement into a hung process
    for holder in (locals().get("trainer"), locals().get("runner")):
        g = getattr(holder, "_graphed", None) or getattr(holder, "graphed", None)
        if g is not None:
            g.release()
    import gc
    gc.collect()
    torch.cuda.synchronize()
    if world > 1:
        bail = threading.Timer(45.0, lambda: os._exit(0))
        bail.daemon = True
        bail.start()
        dist.barrier()
        dist.destroy_process_group()
        bail.cancel()


def dropin_run(cfg, B, dev, steps):
    """INTEGRATION.md route A, timed: deepipr_b200.patch_reference(), then the reference's own model file and
    TrainerPrivate.train (its loop, its F.cross_entropy / accuracy / .item() reads, torch.optim.SGD) drive the CUDA
    path; wrapped in torch.autocast(bf16) so the blocks exchange bf16 activations."""
    import torch
    from deepipr_b200 import _lib as L
    from oracle import ref_bundle
    mods = ref_bundle.import_reference(patched=True)
    ref = ref_bundle.locate()
    name = "alexnet_passport.json" if cfg["net"] == "alexnet" else "resnet18_passport.json"
    pcfg = json.load(open(os.path.join(ref, "passport_configs", name)))
    pkw = mods["experiments.utils"].construct_passport_kwargs_from_dict(
        {"passport_config": pcfg, "norm_type": "bn", "key_type": "random", "sl_ratio": 0.1})
    _seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        if cfg["net"] == "alexnet":
            model = mods["models.alexnet_passport"].AlexNetPassport(3, cfg["classes"], pkw)
        elif cfg["scheme"] == "private":
            model = mods["models.resnet_passport_private"].ResNet18Private(num_classes=cfg["classes"],
                                                                           passport_kwargs=pkw)
        else:
            model = mods["models.resnet_passport"].ResNet18Passport(num_classes=cfg["classes"], passport_kwargs=pkw)
    model = model.to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    tcls = mods["experiments.trainer_private"].TrainerPrivate if cfg["scheme"] == "private" \
        else mods["experiments.trainer"].Trainer
    trainer = tcls(model, opt, None, dev)
    # channels_last batches: the blocks keep the memory format of their input, so this one line on the data side keeps
    # the whole network in NHWC (with NCHW batches the route still works, with a layout conversion around each block)
    data = [(x.contiguous(memory_format=torch.channels_last), t) for x, t in synthetic_batches(cfg, B, 4, device=dev)]
    wm = [(x.contiguous(memory_format=torch.channels_last), t) for x, t in trigger_batches(8, device=dev)] \
        if cfg["trigger"] else None

    def epoch(n):
        with contextlib.redirect_stdout(io.StringIO()), torch.autocast("cuda", dtype=torch.bfloat16):
            return trainer.train(0, [data[i % 4] for i in range(n)], wm)

    epoch(2)
    torch.cuda.synchronize()
    L.load().pp_launch_count(1)
    t0 = time.perf_counter()
    res = epoch(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    per_step = B + (2 if cfg["trigger"] else 0)
    return {"value": steps * per_step / dt, "unit": UNIT, "ms_per_step": dt / steps * 1e3, "steps": steps,
            "library_launches": int(L.load().pp_launch_count(0)), "train_metrics": {k: float(v) for k, v in res.items()},
            "what": "deepipr_b200.patch_reference(); the reference's models/resnet_passport_private.py + "
                    "experiments/trainer_private.py TrainerPrivate.train + torch.optim.SGD, unchanged, under "
                    "torch.autocast(bf16)"}


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the passport-layer training path (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1        # CPU arm: the oracle port of the reference trainer

Workload: one TrainerPrivate-style optimisation step (public + private forward, one backward, SGD) of ResNet-18 with
passport layers in layer4 (passport_configs/resnet18_passport.json) on synthetic CIFAR-10-shaped tensors, bf16
activations, per-GPU batch 1184 = 8 x 148 SMs (weak scaling).  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
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
WORKLOAD = "ResNet18 V2 private-passport (layer4 x5 passport convs) CIFAR10-shaped TrainerPrivate step"
PER_GPU_BATCH = 1184                # 8 images per SM (148 SMs): every conv's tile count is a multiple of the SM count
CPU_BATCH = 64                     # reference default batch (train_v1.py:15); bounded CPU sample
GFLOP_PER_IMAGE_STEP = 6.665       # 2 forwards, 3x fwd FLOPs each


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops_sustained", 1409.4), hbm=d.get("hbm_gbs", 6437.3), src="measured")
    return dict(tflops=1590.0, hbm=6650.0, src="fallback")


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
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_model(num_classes=10, seed=0, blocks=None):
    import contextlib
    import io
    import random
    import numpy as np
    import torch
    from deepipr_b200 import nets
    torch.manual_seed(seed); random.seed(seed); np.random.seed(seed)
    pk = nets.passport_kwargs_from_config(nets.resnet18_passport_config(), "bn", "random", 0.1)
    with contextlib.redirect_stdout(io.StringIO()):
        model = nets.ResNet18("private", num_classes, pk)
    # fixed passports (key_type='random' semantics: U(-1,1), passportconv2d.py:198-207), set up front so the
    # lazily-created-key branch is not part of the timed region
    for m in model.modules():
        if getattr(m, "KIND", None) == "private":
            c = m.conv.in_channels
            h = 8 if m.conv.stride[0] == 2 else 4
            m.set_key(torch.tensor(np.random.uniform(-1, 1, (1, c, h, h)), dtype=torch.float32),
                      torch.tensor(np.random.uniform(-1, 1, (1, c, h, h)), dtype=torch.float32))
    return model


def cpu_reference_run(steps, warmup, batch=CPU_BATCH):
    """The reference's CPU trainer for this path, as restated by the oracle (oracle/passport_oracle.py):
    TrainerPrivate.train step on the box's host cores, all threads."""
    import torch
    from oracle import passport_oracle as po
    torch.set_num_threads(os.cpu_count() or 1)
    model = po.mirror(build_model(), round_bf16=False).train()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, 3, 32, 32, generator=g)
    t = torch.randint(0, 10, (batch,), generator=g)
    for _ in range(warmup):
        po.train_step(model, opt, x, t, private=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        po.train_step(model, opt, x, t, private=True)
    dt = time.perf_counter() - t0
    return dict(value=steps * batch / dt, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"{steps} TrainerPrivate steps at batch {batch} (reference default), fp32, "
                       f"{warmup} warm-up, oracle port of experiments/trainer_private.py:148-177"), dt / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--legs", default="value,e2e,roofline,shared",
                    help="comma list of legs: value (always), e2e, roofline, shared, torch_eager")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    config = {"workload": WORKLOAD, "per_gpu_batch": args.batch, "global_batch": args.batch * max(world, 1),
              "parallelism": f"dp{max(world, 1)}", "passport_layers": "layer4 (5 convs)", "norm": "bn",
              "optimizer": "SGD(0.01, momentum 0.9, wd 1e-4), fused flat step",
              "l2": "per-step working set (~7 GB of activations at batch 1184) >> 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 5))
        cb, sec_per_step = cpu_reference_run(steps, max(1, min(args.warmup, 1)))
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": steps, "warmup": 1, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(config, per_gpu_batch=CPU_BATCH, global_batch=CPU_BATCH, parallelism="cpu"),
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from deepipr_b200 import _lib as L
    from deepipr_b200.parallel import FlatParams, FlatSGD, GradBuckets, broadcast_state
    from deepipr_b200.trainer import StepRunner, accuracy, test_signature

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own logging (NCCL_DEBUG=VERSION/INFO on the box prints
        # "NCCL version ..." there) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()

    model = build_model().to(dev).train()
    broadcast_state(model)
    flat = FlatParams(model.parameters())
    opt = FlatSGD(flat, lr=0.01, momentum=0.9, weight_decay=1e-4)
    buckets = GradBuckets(flat) if world > 1 else None
    runner = StepRunner(model, opt, private=True, buckets=buckets, autocast=True)

    B = args.batch
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    dev_batches = [(torch.randn(B, 3, 32, 32, device=dev, generator=g),
                    torch.randint(0, 10, (B,), device=dev, generator=g)) for _ in range(4)]
    host_batches = [(x.cpu().pin_memory(), t.cpu().pin_memory()) for x, t in dev_batches]

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

    # ---------------- leg 1: inputs resident in HBM
    for i in range(args.warmup):
        runner.step(*dev_batches[i % 4])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.pp_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        runner.step(*dev_batches[i % 4])
    e1.record()
    barrier()
    launches = int(lib.pp_launch_count(0))
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    value = args.steps * B * world / (ms_total * 1e-3)

    legs = set(args.legs.split(","))

    # ---------------- leg 2: end to end through the public step API with host (pinned) buffers
    # Inputs come from pinned host memory every step; the copy of step i+1 is issued on a side stream while step i
    # computes (what DataLoader(pin_memory=True) + .to(device, non_blocking=True) gives the reference loop,
    # trainer_private.py:149-151), so all K copies are inside the timed region but off the critical path.
    copy_stream = torch.cuda.Stream(device=dev)

    def prefetch(i):
        x, t = host_batches[i % 4]
        with torch.cuda.stream(copy_stream):
            xd = x.to(dev, non_blocking=True)
            td = t.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return xd, td, ev

    def e2e_loop(n):
        out = (None,) * 4
        nxt = prefetch(0)
        for i in range(n):
            xd, td, ev = nxt
            main = torch.cuda.current_stream()
            main.wait_event(ev)
            xd.record_stream(main)
            td.record_stream(main)
            if i + 1 < n:
                nxt = prefetch(i + 1)
            loss, sign_loss, preds = runner.step(xd, td)
            # the reference loop's host reads: two accuracies + loss + sign loss (trainer_private.py:163-177)
            out = (accuracy(preds[0], td)[0].item(), accuracy(preds[1], td)[0].item(), sign_loss.item(), loss.item())
        return out

    last = (None,) * 4
    ms_e2e, e2e_value = None, None
    if "e2e" in legs:
        e2e_loop(max(3, args.warmup // 2))
        barrier()
        e0.record()
        last = e2e_loop(args.steps)
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        e2e_value = args.steps * B * world / (ms_e2e * 1e-3)
    h2d = host_batches[0][0].numel() * 4 + host_batches[0][1].numel() * 8
    d2h = 4 * 4

    # ---------------- leg 3: per-kernel roofline of the dominant kernel (CUDA events on its stream)
    roof = None
    roof_w = None
    roof_hbm = None
    if "roofline" in legs:
        # every rank runs the steps (they contain the gradient all-reduce); only rank 0 records kernel events
        if rank == 0:
            lib.pp_profile_enable(1)
        for i in range(3):
            runner.step(*dev_batches[i % 4])
        barrier()
        lib.pp_profile_enable(0)
    if rank == 0 and "roofline" in legs:
        pk = peaks()
        def read(kind, c=0, nout=0, taps=0):
            ms, fl, n = C.c_double(0), C.c_double(0), C.c_int(0)
            lib.pp_profile_read(kind, c, nout, taps, C.byref(ms), C.byref(fl), C.byref(n))
            return ms.value, fl.value, n.value

        (ms0, fl0, n0), (ms1, fl1, n1) = read(0), read(1)
        msp, flp, npl = read(0, 512, 512, 9)      # the 3x3 512->512 passport convs of layer4 (fprop + dgrad)
        if ms0 > 0:
            ach = fl0 / (ms0 * 1e-3) / 1e12
            roof = {"kernel": "tapgemm_kernel (tcgen05 implicit-GEMM conv fprop+dgrad)", "bound": "tensor",
                    "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                    "peak_source": pk["src"] + " bf16 sustained", "traffic": None,
                    "launches_per_step": n0 // 3, "avg_launch_us": ms0 * 1e3 / max(n0, 1),
                    "share_of_step": (ms0 / 3) / (ms_total / args.steps)}
        if msp > 0 and roof is not None:
            ach = flp / (msp * 1e-3) / 1e12
            # traffic: dram__bytes_read+write per launch from the ncu --set full capture of this geometry at batch
            # 1184 (profiles/r1b_ncu_layer4_batch1184.txt: fprop 24.17 + 0.51 MB, dgrad 24.17 MB), scaled to this
            # batch; algorithmic = x + W bytes (the fp32 z / bf16 dx tile stays in the 126 MB L2 for the next pass)
            roof["passport_layer"] = {"geometry": "layer4 3x3 512->512 @4x4, fprop+dgrad launches", "achieved": ach,
                                      "frac": ach / pk["tflops"], "launches_per_step": npl // 3,
                                      "avg_launch_us": msp * 1e3 / max(npl, 1),
                                      "traffic": 24.42e6 * B / 1184.0,
                                      "algorithmic_bytes": 2.0 * (B * 16 * 512 + 512 * 4608)}
        # the HBM-bound passes of the block: algorithmic bytes / CUDA-event time
        hbm_parts, tot_ms, tot_b = {}, 0.0, 0.0
        for kind, name in ((2, "affine_apply (z->y)"), (3, "bwd reduce (dy,z)"), (4, "bwd dz (dy,z->dz)")):
            msk, byk, nk = read(kind)
            if msk > 0:
                hbm_parts[name] = {"achieved": byk / (msk * 1e-3) / 1e9, "launches_per_step": nk // 3,
                                   "avg_launch_us": msk * 1e3 / max(nk, 1),
                                   "share_of_step": (msk / 3) / (ms_total / args.steps)}
                tot_ms += msk
                tot_b += byk
        if tot_ms > 0:
            ach = tot_b / (tot_ms * 1e-3) / 1e9
            roof_hbm = {"kernel": "affine_apply + column_reduce<1> + bwd_dz (norm/affine/ReLU passes)", "bound": "hbm",
                        "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                        "peak_source": pk["src"] + " copy bandwidth",
                        # dram bytes per launch at the layer1 geometry, batch 1184 (profiles/r1b_ncu_layer1_pointwise.txt)
                        "traffic": {"geometry": "layer1 64ch 32x32, batch 1184",
                                    "affine_apply": 421.5e6, "bwd_reduce": 469.4e6, "bwd_dz": 599.1e6,
                                    "algorithmic": {"affine_apply": 465.6e6, "bwd_reduce": 465.6e6,
                                                    "bwd_dz": 620.8e6}},
                        "share_of_step": (tot_ms / 3) / (ms_total / args.steps), "per_kernel": hbm_parts,
                        "note": "small layers (layer3/4) re-read z/dy from the 126 MB L2, so a per-kernel figure can "
                                "exceed the HBM peak; the aggregate is dominated by layer1/2"}
        if ms1 > 0:
            ach = fl1 / (ms1 * 1e-3) / 1e12
            roof_w = {"kernel": "wgrad_kernel (tcgen05, MN-major)", "bound": "tensor", "achieved": ach,
                      "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                      "launches_per_step": n1 // 3, "avg_launch_us": ms1 * 1e3 / max(n1, 1),
                      "share_of_step": (ms1 / 3) / (ms_total / args.steps)}

    # ---------------- optional leg: the same step in stock PyTorch eager (cuDNN/cuBLAS, autocast bf16,
    # channels_last) on this GPU — the oracle mirror moved to the device.  Informational only ("the honest bar",
    # BASELINE.md §3a); not part of the default run.
    torch_eager = None
    if rank == 0 and "torch_eager" in legs:
        import torch.nn.functional as TF
        from oracle import passport_oracle as po
        torch.backends.cudnn.benchmark = True                      # train_v1.py:8 / train_v23.py:8
        ref = po.mirror(build_model(), round_bf16=False).to(dev).to(memory_format=torch.channels_last).train()
        ropt = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
        rlosses = po.sign_loss_modules(ref)

        def ref_step(x, t):
            ropt.zero_grad()
            for m in rlosses:
                m.reset()
            loss = torch.zeros((), device=dev)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                for ind in range(2):
                    loss = loss + TF.cross_entropy(ref(x, ind=ind).float(), t)
            sl = torch.zeros((), device=dev)
            for m in rlosses:
                sl = sl + m.loss
            (loss + sl).backward()
            ropt.step()

        xs = [(x.contiguous(memory_format=torch.channels_last), t) for x, t in dev_batches]
        for i in range(max(3, args.warmup)):
            ref_step(*xs[i % 4])
        torch.cuda.synchronize()
        e0.record()
        for i in range(args.steps):
            ref_step(*xs[i % 4])
        e1.record()
        torch.cuda.synchronize()
        ms_ref = e0.elapsed_time(e1)
        torch_eager = {"value": args.steps * B / (ms_ref * 1e-3), "unit": UNIT, "ms_per_step": ms_ref / args.steps,
                       "what": "oracle mirror on cuda: torch eager + cuDNN, autocast bf16, channels_last, "
                               "cudnn.benchmark, torch.optim.SGD(foreach); same batch, inputs resident"}
        del ref, ropt

    # ---------------- optional extra: the same step with the public/private passes sharing the passport-free
    # trunk (nets.ResNet18.share_trunk, identical results up to summation order).  Reported separately; `value`
    # above always executes both full passes like the reference does.
    shared = None
    if "shared" in legs:
        model.share_trunk = True
        for i in range(3):
            runner.step(*dev_batches[i % 4])
        barrier()
        e0.record()
        for i in range(args.steps):
            runner.step(*dev_batches[i % 4])
        e1.record()
        barrier()
        ms_sh = max_over_ranks(e0.elapsed_time(e1))
        shared = {"value": args.steps * B * world / (ms_sh * 1e-3), "unit": UNIT, "ms_per_step": ms_sh / args.steps,
                  "what": "public+private passes share the stem..layer3 forward/backward (common-subexpression "
                          "elimination across the two model calls of trainer_private.py:159-161)"}
        model.share_trunk = False

    sig = test_signature(model) if rank == 0 else {}
    model.train()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and "e2e" in legs:
        cpu_baseline, _ = cpu_reference_run(args.cpu_steps, 1)

    if rank == 0:
        pk = peaks()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": (ms_e2e / args.steps) if ms_e2e else None},
                "gpu_launches": launches,
                "roofline": roof, "roofline_wgrad": roof_w, "roofline_hbm": roof_hbm,
                "conv_roofline_frac_whole_step": (value / world) * GFLOP_PER_IMAGE_STEP * 1e9 / (pk["tflops"] * 1e12),
                "cpu_baseline": cpu_baseline, "torch_eager_gpu": torch_eager, "value_shared_trunk": shared,
                "sign_bit_accuracy": (sum(sig.values()) / len(sig)) if sig else None,
                "last_step": {"acc_public": last[0], "acc_private": last[1], "sign_loss": last[2], "loss": last[3]}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

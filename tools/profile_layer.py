"""Micro-benchmark / profiling target: forward + backward of ONE passport block at a BASELINE geometry.

    python tools/profile_layer.py --layer layer4 --batch 1024 --iters 20
    ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -c 2 -o gpurun_out/prof \
        python tools/profile_layer.py --layer layer4 --iters 1 --warmup 1

Prints CUDA-event timings of the whole block step and, via the library's per-kernel event instrumentation, the
achieved TFLOP/s of the two tensor-core kernels.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LAYERS = {  # name: (C, O, k, s, p, H)
    "layer1": (64, 64, 3, 1, 1, 32), "layer2": (128, 128, 3, 1, 1, 16), "layer3": (256, 256, 3, 1, 1, 8),
    "layer4": (512, 512, 3, 1, 1, 4), "layer4s2": (256, 512, 3, 2, 1, 8), "layer4sc": (256, 512, 1, 2, 0, 8),
    "stem": (3, 64, 3, 1, 1, 32), "alex4": (192, 384, 3, 1, 1, 8), "alex5": (384, 256, 3, 1, 1, 8), "imagenet4": (512, 512, 3, 1, 1, 7),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layer", default="layer4")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--kind", default="private", choices=["private", "v1", "conv"])
    ap.add_argument("--no-fused", action="store_true", help="kernel sequence instead of the single-kernel block")
    args = ap.parse_args()
    import torch
    from deepipr_b200 import _lib as L
    from deepipr_b200 import layers
    from tests.helpers import quiet, seed_all
    Ci, O, k, s, p, H = LAYERS[args.layer]
    seed_all(0)
    kw = {"norm_type": "bn", "key_type": "random", "sign_loss": 0.1}
    if args.kind == "private":
        m = quiet(layers.PassportPrivateBlock, Ci, O, k, s, p, kw)
    elif args.kind == "v1":
        m = quiet(layers.PassportBlock, Ci, O, k, s, p, kw)
    else:
        m = layers.ConvBlock(Ci, O, k, s, p, bn="bn")
    if args.kind != "conv":
        m.set_key(torch.rand(1, Ci, H, H) * 2 - 1, torch.rand(1, Ci, H, H) * 2 - 1)
    m = m.cuda().train()
    x = torch.randn(args.batch, Ci, H, H, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x.requires_grad_(Ci != 3)      # the network input carries no gradient
    P = (H + 2 * p - k) // s + 1
    gy = torch.randn(args.batch, O, P, P, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    lib = L.load()
    if args.no_fused:
        lib.pp_debug_fused(0)

    def step():
        if args.kind == "private":
            y = m(x, False, 1)
        elif args.kind == "v1":
            y = m(x)
        else:
            y = m(x)
        y.backward(gy)
        m.zero_grad(set_to_none=True)
        x.grad = None

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    lib.pp_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    lib.pp_profile_enable(0)
    out = {"layer": args.layer, "batch": args.batch, "kind": args.kind, "ms_per_block_step": e0.elapsed_time(e1) / args.iters}
    flops = 2.0 * args.batch * P * P * O * Ci * k * k
    out["block_tflops_3x"] = 3 * flops / (out["ms_per_block_step"] * 1e-3) / 1e12
    for kind, name in ((0, "tapgemm"), (1, "wgrad"), (5, "passport_fused")):
        ms, fl, n = C.c_double(0), C.c_double(0), C.c_int(0)
        lib.pp_profile_read(kind, 0, 0, 0, C.byref(ms), C.byref(fl), C.byref(n))
        if n.value:
            out[name] = {"launches": n.value, "avg_us": ms.value * 1e3 / n.value, "tflops": fl.value / (ms.value * 1e-3) / 1e12}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

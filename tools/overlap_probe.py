"""Can a tensor-core weight-gradient kernel and an HBM-bound streaming pass share the GPU?  Times wgrad alone, a streaming
kernel alone (pp_add_relu_fwd over a tensor of the layer's size: 6 bytes / element) and both on two streams."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from deepipr_b200 import _lib as L
from deepipr_b200 import functional as F_

LAYERS = {"layer1": (64, 64, 32), "layer2": (128, 128, 16), "layer3": (256, 256, 8), "layer4": (512, 512, 4)}


def main():
    lib = L.load()
    N = 1026
    for name, (Ci, O, H) in LAYERS.items():
        spec = F_.ConvSpec(Ci, O, 3, 3, 1, 1)
        x = torch.randn(N, Ci, H, H, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        dz = torch.randn(N, H, H, O, device="cuda").to(torch.bfloat16)
        a = torch.randn(N, O, H, H, device="cuda").to(torch.bfloat16)
        b = torch.randn_like(a)
        y = torch.empty_like(a)
        # a larger streaming workload too (the passes of the NEXT layer in backward order are as large or larger)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def wgrad():
            F_.conv_wgrad(dz, x, spec)

        def stream_pass():
            lib.pp_add_relu_fwd(C.c_size_t(a.numel()), L.ptr(a), L.ptr(b), L.ptr(y),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))

        def timed(fa, fb, iters=20):
            for _ in range(3):
                if fa:
                    with torch.cuda.stream(s1):
                        fa()
                if fb:
                    with torch.cuda.stream(s2):
                        fb()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
            for _ in range(iters):
                if fa:
                    with torch.cuda.stream(s1):
                        fa()
                if fb:
                    with torch.cuda.stream(s2):
                        fb()
            torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters * 1e3

        tw, ts, tb = timed(wgrad, None), timed(None, stream_pass), timed(wgrad, stream_pass)
        print(f"{name}: wgrad {tw:.1f} us, streaming pass {ts:.1f} us, both {tb:.1f} us, sum {tw + ts:.1f} us, "
              f"hidden {tw + ts - tb:.1f} us = {100 * (tw + ts - tb) / min(tw, ts):.0f} % of the shorter one")


if __name__ == "__main__":
    main()

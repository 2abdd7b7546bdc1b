"""Kernel bring-up diagnostics for the GPU box.

    python tools/diag_kernels.py            # runs every case group in its own subprocess (a device-side trap
                                            # in one group cannot poison the others), writes gpurun_out/diag.json
    python tools/diag_kernels.py --group fwd_tc

References here are torch GPU ops in fp32 with TF32 off (fast, independent of our kernels) — this is a
debugging tool, not a parity test (those live in tests/ and use the CPU oracle).
"""
import argparse
import json
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["pointwise", "fwd_simt", "fwd_tc", "dgrad", "wgrad", "block"]


def _setup():
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch


def bf16r(t):
    import torch
    return t.to(torch.bfloat16).to(torch.float32)


def err_stats(got, ref):
    got = got.double()
    ref = ref.double()
    diff = (got - ref).abs()
    denom = ref.norm().item() or 1.0
    out = dict(rel_l2=(got - ref).norm().item() / denom, max_abs=diff.max().item(), ref_max=ref.abs().max().item())
    if diff.numel():
        flat = diff.flatten()
        idx = int(flat.argmax())
        out["argmax"] = idx
        bad = (diff > 1e-2 * max(out["ref_max"], 1e-6))
        out["bad_frac"] = bad.double().mean().item()
        if bad.any() and bad.dim() >= 2:
            rows = bad.reshape(-1, bad.shape[-1]).any(dim=1)
            cols = bad.reshape(-1, bad.shape[-1]).any(dim=0)
            out["bad_rows"] = int(rows.sum())
            out["bad_cols"] = int(cols.sum())
            out["first_bad_rows"] = [int(i) for i in rows.nonzero().flatten()[:8]]
            out["first_bad_cols"] = [int(i) for i in cols.nonzero().flatten()[:8]]
    return out


CONV_GEOMS = [
    # name, N, C, H, O, k, s, p
    ("1x1_c64_o64_onetile", 2, 64, 8, 64, 1, 1, 0),
    ("1x1_c128_o128", 2, 128, 8, 128, 1, 1, 0),
    ("3x3_c64_o64_8x8", 2, 64, 8, 64, 3, 1, 1),
    ("3x3_c64_o64_tail48", 3, 64, 4, 64, 3, 1, 1),
    ("3x3_c64_o128_tail144", 9, 64, 4, 128, 3, 1, 1),
    ("3x3_c192_o384_alex", 4, 192, 8, 384, 3, 1, 1),
    ("3x3_c512_o512_layer4", 64, 512, 4, 512, 3, 1, 1),
    ("3x3_s2_c256_o512", 32, 256, 8, 512, 3, 2, 1),
    ("1x1_s2_c256_o512", 32, 256, 8, 512, 1, 2, 0),
    ("3x3_c64_o64_32x32_layer1", 4, 64, 32, 64, 3, 1, 1),
    ("3x3_s2_c64_o128_32x32", 4, 64, 32, 128, 3, 2, 1),
    ("3x3_c512_o512_many_tiles", 512, 512, 4, 512, 3, 1, 1),
    ("3x3_c256_o256_7x7_odd", 5, 256, 7, 256, 3, 1, 1),
    ("3x3_s2_c128_o256_14x14", 3, 128, 14, 256, 3, 2, 1),
]


def make_case(torch, N, C, H, O, k, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = bf16r(torch.randn(N, C, H, H, generator=g)).cuda()
    w = bf16r(torch.randn(O, C, k, k, generator=g) * (2.0 / (C * k * k)) ** 0.5).cuda()
    return x, w


def group_pointwise(res):
    torch = _setup()
    from deepipr_b200 import functional as F_
    from deepipr_b200 import _lib as L
    import ctypes as C
    spec = F_.ConvSpec(64, 128, 3, 3, 1, 1)
    x, w = make_case(torch, 2, 64, 8, 128, 3)
    prep = F_.prepare_weight(w, spec, True)
    res["weight_prep_wf"] = err_stats(prep.wf.float(), w.permute(0, 2, 3, 1).contiguous())
    res["weight_prep_wd"] = err_stats(prep.wd.float(), w.permute(1, 2, 3, 0).contiguous())
    # key pool + affine
    for (s, p, k, Bk) in ((1, 1, 3, 1), (2, 1, 3, 2), (2, 0, 1, 1)):
        spec = F_.ConvSpec(64, 128, k, k, s, p)
        _, w = make_case(torch, 1, 64, 8, 128, k, seed=1)
        key = bf16r(torch.rand(Bk, 64, 8, 8) * 2 - 1).cuda()
        skey = bf16r(torch.rand(Bk, 64, 8, 8) * 2 - 1).cuda()
        prep = F_.prepare_weight(w, spec, False)
        Ss, Sk = F_.key_pool(skey, spec), F_.key_pool(key, spec)
        b = torch.sign(torch.rand(128) - 0.5).cuda()
        actx = F_.AffineCtx(spec, prep, Ss, Sk, b, 0.1)
        wreq = w.clone().requires_grad_(True)
        gamma, beta, loss, acc = F_.passport_affine(wreq, actx)
        ref_g = torch.nn.functional.conv2d(skey.double(), w.double(), None, s, p).mean(dim=(0, 2, 3))
        ref_b = torch.nn.functional.conv2d(key.double(), w.double(), None, s, p).mean(dim=(0, 2, 3))
        tag = f"affine_k{k}s{s}B{Bk}"
        res[tag + "_gamma"] = err_stats(gamma, ref_g)
        res[tag + "_beta"] = err_stats(beta, ref_b)
        ref_loss = (0.1 * torch.relu(-b.double() * ref_g + 0.1)).sum() + 1e-5 * ref_g.pow(2).sum()
        res[tag + "_loss"] = dict(got=float(loss), ref=float(ref_loss))
        res[tag + "_acc"] = dict(got=float(acc), ref=float((torch.sign(b.double()) == torch.sign(ref_g)).double().mean()))
        # backward vs autograd of the torch formulation
        r1, r2 = torch.randn(128).cuda(), torch.randn(128).cuda()
        ((gamma * r1).sum() + (beta * r2).sum() + loss).backward()
        w2 = w.clone().double().requires_grad_(True)
        g2 = torch.nn.functional.conv2d(skey.double(), w2, None, s, p).mean(dim=(0, 2, 3))
        b2 = torch.nn.functional.conv2d(key.double(), w2, None, s, p).mean(dim=(0, 2, 3))
        l2 = (0.1 * torch.relu(-b.double() * g2 + 0.1)).sum() + 1e-5 * g2.pow(2).sum()
        ((g2 * r1.double()).sum() + (b2 * r2.double()).sum() + l2).backward()
        res[tag + "_dw"] = err_stats(wreq.grad, w2.grad)
    # sgd
    p = torch.randn(10007).cuda(); g = torch.randn(10007).cuda(); buf = torch.zeros(10007).cuda()
    pr = p.clone();
    lib = L.load()
    for step in range(3):
        L.check(lib.pp_sgd_step(C.c_size_t(p.numel()), L.ptr(p), L.ptr(g), L.ptr(buf), 0.1, 0.9, 1e-4, int(step == 0),
                                None))
    pref = torch.nn.Parameter(pr.clone()); opt = torch.optim.SGD([pref], lr=0.1, momentum=0.9, weight_decay=1e-4)
    for step in range(3):
        pref.grad = g.clone(); opt.step()
    res["sgd"] = err_stats(p, pref.detach())


def _conv_fwd_cases(res, algo, geoms):
    torch = _setup()
    from deepipr_b200 import functional as F_
    for (name, N, C, H, O, k, s, p) in geoms:
        try:
            spec = F_.ConvSpec(C, O, k, k, s, p)
            x, w = make_case(torch, N, C, H, O, k)
            prep = F_.prepare_weight(w, spec, False)
            t0 = time.time()
            z = F_.conv_fwd_raw(x, prep, spec, z_f32=True, algo=algo)
            torch.cuda.synchronize()
            ref = torch.nn.functional.conv2d(x, w, None, s, p).permute(0, 2, 3, 1).contiguous()
            st = err_stats(z, ref)
            st["ms_first_call"] = (time.time() - t0) * 1e3
            zb = F_.conv_fwd_raw(x, prep, spec, z_f32=False, algo=algo)
            st["bf16_rel_l2"] = err_stats(zb.float(), ref)["rel_l2"]
            res[name] = st
        except Exception as e:  # noqa
            res[name] = dict(error=str(e)[:400])
            if "CUDA" in str(e) or "launch" in str(e):
                raise


def group_fwd_simt(res):
    from deepipr_b200 import _lib as L
    geoms = [g for g in CONV_GEOMS if g[1] * g[3] * g[3] <= 4096][:8] + [("stem_c3", 8, 3, 32, 64, 3, 1, 1)]
    _conv_fwd_cases(res, L.PP_ALGO_SIMT, geoms)


def group_fwd_tc(res):
    from deepipr_b200 import _lib as L
    _conv_fwd_cases(res, L.PP_ALGO_TCGEN05, CONV_GEOMS)


def group_dgrad(res):
    torch = _setup()
    from deepipr_b200 import functional as F_
    from deepipr_b200 import _lib as L
    for (name, N, C, H, O, k, s, p) in CONV_GEOMS:
        for algo, tag in ((L.PP_ALGO_SIMT, "simt"), (L.PP_ALGO_TCGEN05, "tc")):
            if tag == "simt" and N * H * H > 4096:
                continue
            try:
                spec = F_.ConvSpec(C, O, k, k, s, p)
                x, w = make_case(torch, N, C, H, O, k)
                P, Q = spec.out_hw(H, H)
                dz = bf16r(torch.randn(N, P, Q, O)).cuda()
                prep = F_.prepare_weight(w, spec, True)
                dx = F_.conv_dgrad(dz, prep, spec, N, H, H, algo=algo)
                torch.cuda.synchronize()
                ref = torch.nn.grad.conv2d_input((N, C, H, H), w, dz.permute(0, 3, 1, 2), s, p)
                res[f"{name}_{tag}"] = err_stats(dx.float(), ref.permute(0, 2, 3, 1).contiguous())
            except Exception as e:  # noqa
                res[f"{name}_{tag}"] = dict(error=str(e)[:400])
                if "CUDA" in str(e) or "launch" in str(e):
                    raise


def group_wgrad(res):
    torch = _setup()
    from deepipr_b200 import functional as F_
    from deepipr_b200 import _lib as L
    geoms = CONV_GEOMS + [("stem_c3", 8, 3, 32, 64, 3, 1, 1)]
    for (name, N, C, H, O, k, s, p) in geoms:
        for algo, tag in ((L.PP_ALGO_SIMT, "simt"), (L.PP_ALGO_TCGEN05, "tc")):
            if tag == "simt" and N * H * H > 8192:
                continue
            if tag == "tc" and C % 64:
                continue
            try:
                spec = F_.ConvSpec(C, O, k, k, s, p)
                x, w = make_case(torch, N, C, H, O, k)
                P, Q = spec.out_hw(H, H)
                dz = bf16r(torch.randn(N, P, Q, O)).cuda()
                dw = F_.conv_wgrad(dz, x, spec, algo=algo)
                torch.cuda.synchronize()
                ref = torch.nn.grad.conv2d_weight(x, (O, C, k, k), dz.permute(0, 3, 1, 2), s, p)
                res[f"{name}_{tag}"] = err_stats(dw, ref)
            except Exception as e:  # noqa
                res[f"{name}_{tag}"] = dict(error=str(e)[:400])
                if "CUDA" in str(e) or "launch" in str(e):
                    raise


def group_block(res):
    torch = _setup()
    from deepipr_b200 import layers
    from oracle import passport_oracle as po
    from tests.helpers import quiet, seed_all
    cases = [
        ("v1_bn_c512", "v1", 512, 512, 3, 1, 1, "bn", 32, 4),
        ("v1_none_c256_s2", "v1", 256, 512, 3, 2, 1, "none", 16, 8),
        ("private_bn_c512", "private", 512, 512, 3, 1, 1, "bn", 32, 4),
        ("conv_bn_c64", "conv", 64, 64, 3, 1, 1, "bn", 8, 32),
        ("conv_bn_stem", "conv", 3, 64, 3, 1, 1, "bn", 8, 32),
    ]
    for (name, kind, i, o, ks, s, pd, norm, N, H) in cases:
        try:
            seed_all(0)
            kw = {"norm_type": norm, "key_type": "random", "sign_loss": 0.1}
            if kind == "v1":
                m = quiet(layers.PassportBlock, i, o, ks, s, pd, kw)
            elif kind == "private":
                m = quiet(layers.PassportPrivateBlock, i, o, ks, s, pd, kw)
            else:
                m = layers.ConvBlock(i, o, ks, s, pd, bn=norm)
            with torch.no_grad():
                m.conv.weight.copy_(bf16r(m.conv.weight))
            if kind != "conv":
                m.set_key(bf16r(torch.rand(1, i, H, H) * 2 - 1), bf16r(torch.rand(1, i, H, H) * 2 - 1))
            x = bf16r(torch.randn(N, i, H, H))
            orc = po.mirror(m, round_bf16=True)
            m = m.cuda()
            inds = (0, 1) if kind == "private" else (0,)
            outs = {}
            for tag, mod, dev in (("gpu", m, "cuda"), ("ref", orc, "cpu")):
                mod.train()
                xx = x.to(dev).clone().requires_grad_(True)
                for sl in mod.modules():
                    if hasattr(sl, "scale_cache"):
                        sl.reset()
                tot = 0
                ys = []
                torch.manual_seed(5)
                for ind in inds:
                    y = mod(xx, False, ind) if kind == "private" else (mod(xx) if kind == "conv" else mod(xx, False))
                    r = bf16r(torch.randn(y.shape)).to(dev)
                    tot = tot + (y.float() * r).sum()
                    ys.append(y.detach().float().cpu())
                sl_tot = 0
                for sl in mod.modules():
                    if hasattr(sl, "scale_cache"):
                        sl_tot = sl_tot + sl.loss
                (tot + sl_tot).backward()
                outs[tag] = dict(y=ys, dx=xx.grad.detach().float().cpu(), sl=float(sl_tot),
                                 grads={k: p.grad.detach().float().cpu() for k, p in mod.named_parameters()
                                        if p.grad is not None})
            r = {}
            for kidx in range(len(inds)):
                r[f"y{kidx}"] = err_stats(outs["gpu"]["y"][kidx], outs["ref"]["y"][kidx])["rel_l2"]
            r["dx"] = err_stats(outs["gpu"]["dx"], outs["ref"]["dx"])["rel_l2"]
            r["sign_loss"] = (outs["gpu"]["sl"], outs["ref"]["sl"])
            for kname, gref in outs["ref"]["grads"].items():
                gk = outs["gpu"]["grads"].get(kname)
                if gk is None and kname == "conv.weight":
                    gk = outs["gpu"]["grads"].get("weight")
                if gk is None and kname == "weight":
                    gk = outs["gpu"]["grads"].get("conv.weight")
                r["d_" + kname] = err_stats(gk, gref)["rel_l2"] if gk is not None else "missing"
            res[name] = r
        except Exception as e:  # noqa
            res[name] = dict(error=traceback.format_exc()[-600:])
            if "CUDA" in str(e) or "launch" in str(e):
                raise


def run_group(name):
    res = {}
    try:
        globals()["group_" + name](res)
    except Exception:
        res["__exception__"] = traceback.format_exc()[-1500:]
        try:
            from deepipr_b200 import _lib as L
            res["__last_error__"] = L.last_error()
        except Exception:
            pass
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "diag.json"))
    args = ap.parse_args()
    if args.group:
        print(json.dumps(run_group(args.group)))
        return
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    allres = {}
    for g in GROUPS:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", g], capture_output=True,
                               text=True, timeout=420)
            line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else ""
            try:
                allres[g] = json.loads(line)
            except Exception:
                allres[g] = dict(__raw_stdout__=p.stdout[-1500:], __stderr__=p.stderr[-2500:], rc=p.returncode)
            if p.returncode != 0 and isinstance(allres[g], dict):
                allres[g]["__stderr__"] = p.stderr[-2500:]
        except subprocess.TimeoutExpired as e:
            allres[g] = dict(__timeout__=True, out=str(e.stdout)[-800:] if e.stdout else "")
        allres[g]["__seconds__"] = round(time.time() - t0, 1)
        with open(args.out, "w") as f:
            json.dump(allres, f, indent=1)
    # compact console summary
    for g, r in allres.items():
        print("=====", g, r.get("__seconds__"))
        for k, v in r.items():
            if k.startswith("__"):
                print("  ", k, str(v)[-1200:])
            elif isinstance(v, dict) and "rel_l2" in v:
                print(f"   {k:38s} rel_l2={v['rel_l2']:.3e} max_abs={v['max_abs']:.3e} bad={v.get('bad_frac', 0):.3f}"
                      + (f" rows={v.get('bad_rows')} cols={v.get('bad_cols')} r0={v.get('first_bad_rows')} c0={v.get('first_bad_cols')}" if v.get("bad_frac", 0) > 0 else "")
                      + (f" bf16={v['bf16_rel_l2']:.2e}" if "bf16_rel_l2" in v else ""))
            else:
                print("  ", k, v)


if __name__ == "__main__":
    main()

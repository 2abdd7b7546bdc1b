"""Run one of the reference's OWN entry scripts (train_v1.py / train_v23.py) unchanged — TEST INFRASTRUCTURE ONLY.

    python -m oracle.run_reference --script train_v23.py --flavour patched|stock --device cuda|cpu --out res.json \
        [--train-batches 2 --val-batches 1 --wm-batches 1 --seed 0 --pretrained] -- <the script's own arguments>

`--flavour stock` executes the unmodified reference (torch eager).  `--flavour patched` calls
deepipr_b200.patch_reference() first, so the same script, experiment class, model files and trainer loop run on this
repository's blocks (INTEGRATION.md route A).  Nothing of the reference is edited; the harness only

  * replaces dataset.prepare_dataset / prepare_wm by synthetic CIFAR-shaped loaders (no datasets, no network here),
  * with --pretrained writes a randomly initialised "pretrained" normal network (the reference would download
    torchvision weights) so that --key-type shuffle / image exercise passport_generator.set_key,
  * for --device cpu swaps the hard-coded torch.device('cuda') of experiments/base.py:29.

The metrics the reference wrote to logs/<run>/history.csv are returned as JSON, plus checksums of the final model.
"""
import argparse
import contextlib
import csv
import glob
import io
import json
import os
import random
import runpy
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _synthetic_loaders(args, script_args):
    import torch
    from torch.utils.data import DataLoader, TensorDataset

    def opt(name, default, cast=int):
        return cast(script_args[script_args.index(name) + 1]) if name in script_args else default

    bs = opt("--batch-size", 64)
    ds = opt("--dataset", "cifar10", str)
    classes = {"cifar10": 10, "cifar100": 100, "imagenet1000": 1000}[ds]
    size = 224 if ds == "imagenet1000" else 32
    g = torch.Generator().manual_seed(1234 + args.seed)

    def make(nb, batch, shuffle=False, ncls=classes):
        x = torch.randn(nb * batch, 3, size, size, generator=g)
        t = torch.randint(0, ncls, (nb * batch,), generator=g)
        return DataLoader(TensorDataset(x, t), batch_size=batch, shuffle=shuffle, drop_last=True)

    # the validation set doubles as the passport candidate pool: passport_generator.get_key samples 20 images from it
    train, val = make(args.train_batches, bs), make(max(args.val_batches, 1, -(-20 // (bs * 2))), bs * 2)
    wm = make(args.wm_batches, 2, ncls=10)                 # prepare_wm: batch 2, CIFAR labels (dataset.py:168-193)
    return train, val, wm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--script", default="train_v23.py")
    ap.add_argument("--flavour", choices=["patched", "stock"], required=True)
    ap.add_argument("--device", choices=["cuda", "cpu"], default="cuda")
    ap.add_argument("--out", required=True)
    ap.add_argument("--train-batches", type=int, default=2)
    ap.add_argument("--val-batches", type=int, default=1)
    ap.add_argument("--wm-batches", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--pretrained", action="store_true")
    ap.add_argument("--autocast", action="store_true", help="run the script under torch.autocast(bf16)")
    ap.add_argument("--save-model", default=None, help="copy the run's models/last.pth (a plain state_dict) here")
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    args = ap.parse_args()
    script_args = [a for a in args.rest if a != "--"]

    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "0")       # trainers wrap nn.DataParallel when they see >1 device
    from oracle import ref_bundle
    work = ref_bundle.workdir()
    try:
        os.chdir(work)
        sys.path.insert(0, work)
        if args.flavour == "patched":
            import deepipr_b200
            deepipr_b200.patch_reference()
        import numpy as np
        import torch
        torch.manual_seed(args.seed); random.seed(args.seed); np.random.seed(args.seed)

        import dataset as ref_dataset
        train, val, wm = _synthetic_loaders(args, script_args)
        ref_dataset.prepare_dataset = lambda a: (train, val)
        ref_dataset.prepare_wm = lambda *a, **k: wm
        if args.device == "cpu":
            import experiments.base as eb

            class _TorchCpu:
                def __getattr__(self, name):
                    return getattr(torch, name)

                @staticmethod
                def device(*_a, **_k):
                    return torch.device("cpu")
            eb.torch = _TorchCpu()
        if args.pretrained:
            arch = script_args[script_args.index("--arch") + 1] if "--arch" in script_args else "alexnet"
            ncls = 100 if "cifar100" in script_args else 10
            norm = script_args[script_args.index("--norm-type") + 1] if "--norm-type" in script_args else "bn"
            if arch == "resnet":
                from models.resnet_normal import ResNet18
                pre = ResNet18(num_classes=ncls, norm_type=norm)
            else:
                from models.alexnet_normal import AlexNetNormal
                pre = AlexNetNormal(3, ncls, norm_type=norm)
            torch.save(pre.state_dict(), os.path.join(work, "pretrained.pth"))
            script_args += ["--pretrained-path", os.path.join(work, "pretrained.pth")]
            torch.manual_seed(args.seed); random.seed(args.seed); np.random.seed(args.seed)

        sys.argv = [args.script] + script_args
        log = io.StringIO()
        ctx = torch.autocast(args.device, dtype=torch.bfloat16) if args.autocast else contextlib.nullcontext()
        with contextlib.redirect_stdout(log), ctx:
            runpy.run_path(os.path.join(work, args.script), run_name="__main__")

        hist = sorted(glob.glob(os.path.join(work, "logs", "*", "*", "history.csv")))
        assert hist, "the reference wrote no history.csv:\n" + log.getvalue()[-2000:]
        with open(hist[-1]) as f:
            rows = list(csv.DictReader(f, delimiter=",", quotechar="'"))
        run_dir = os.path.dirname(hist[-1])
        sd = torch.load(os.path.join(run_dir, "models", "last.pth"), map_location="cpu")
        if args.save_model:
            shutil.copyfile(os.path.join(run_dir, "models", "last.pth"), args.save_model)
        out = {"history": [{k: float(v) for k, v in r.items()} for r in rows],
               "state_keys": sorted(sd.keys()),
               "param_abs_sums": {k: v.double().abs().sum().item() for k, v in sd.items()
                                  if v.dtype.is_floating_point},
               "flavour": args.flavour, "device": args.device, "script": args.script, "args": script_args,
               "log_tail": log.getvalue()[-1500:]}
        if args.flavour == "patched" and args.device == "cuda":
            from deepipr_b200 import _lib
            out["library_launches"] = int(_lib.load().pp_launch_count(0))
        with open(args.out, "w") as f:
            json.dump(out, f)
    finally:
        os.chdir(ROOT)
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()

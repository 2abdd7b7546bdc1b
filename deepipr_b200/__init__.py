"""deepipr_b200 — B200-native (sm_100a) passport-layer training path behind the DeepIPR module surface.

  layers      PassportBlock / PassportPrivateBlock / ConvBlock / SignLoss   (drop-in for the reference's
              models/layers/*.py and models/losses/sign_loss.py)
  functional  autograd operators over the C ABI of libpassport_sm100.so (include/passport_sm100.h)
  nets        ResNet-18 / AlexNet wiring with the reference's state_dict keys
  trainer     V1 and V2/V3 loops with the reference's step semantics, DDP instead of DataParallel
  parallel    flat gradient buckets + NCCL all-reduce, fused flat SGD
"""
import sys
import types

__all__ = ["patch_reference", "layers", "functional", "nets", "trainer", "parallel"]


def patch_reference(conv_block=True, precision=None):
    """Make the reference's own import paths resolve to this package's blocks, so that its models/,
    experiments/ and train_v1.py / train_v23.py run unchanged (call BEFORE importing reference code):

        import deepipr_b200; deepipr_b200.patch_reference()
        sys.path.insert(0, "/path/to/DeepIPR"); runpy.run_path("train_v23.py", run_name="__main__")

    conv_block=False leaves models.layers.conv2d.ConvBlock to the reference (needed for its CPU-only
    plumbing runs such as AlexNet-normal on CPU; the fused blocks here have no CPU path).
    precision='tf32' runs fp32 inputs through the tcgen05 kind::tf32 kernels with fp32 activations (what the
    reference's fp32 scripts get from cuDNN: train_v1.py / train_v23.py without autocast); the default 'bf16' converts
    block inputs to bf16 (layers.set_precision).
    """
    from . import layers
    if precision is not None:
        layers.set_precision(precision)

    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__deepipr_b200__ = True
        sys.modules[name] = m
        return m

    module("models.layers.passportconv2d", PassportBlock=layers.PassportBlock, SignLoss=layers.SignLoss)
    module("models.layers.passportconv2d_private", PassportPrivateBlock=layers.PassportPrivateBlock,
           SignLoss=layers.SignLoss)
    module("models.losses.sign_loss", SignLoss=layers.SignLoss)
    if conv_block:
        module("models.layers.conv2d", ConvBlock=layers.ConvBlock)

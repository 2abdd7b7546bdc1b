"""CPU (fp64) checks of the algebra the CUDA kernels rely on — the identities are restated with plain torch ops and
compared with autograd of the operators the reference calls.  No library code runs here; the GPU parity tests
(tests/test_gpu_parity.py) check the kernels themselves."""
import torch
import torch.nn.functional as F


def test_group_norm_backward_coefficients():
    """groupnorm.cu: y = relu(a[n,o] z + b[n,o]) and dz = k1[n,o] dy_m + k2[n,o] z + k3[n,o] with the coefficient
    formulas of gn_fwd_coef_kernel / gn_bwd_coef_kernel, dgamma / dbeta as in gn_dparam_kernel — against autograd of
    F.group_norm (passportconv2d.py:59-60) followed by the passport affine and ReLU (:219-222)."""
    torch.manual_seed(0)
    N, O, H, G, eps = 3, 32, 5, 2, 1e-5
    z = torch.randn(N, O, H, H, dtype=torch.double, requires_grad=True)
    gamma = torch.randn(O, dtype=torch.double, requires_grad=True)
    beta = torch.randn(O, dtype=torch.double, requires_grad=True)
    y = torch.relu(gamma.view(1, -1, 1, 1) * F.group_norm(z, G, None, None, eps) + beta.view(1, -1, 1, 1))
    dy = torch.randn_like(y)
    y.backward(dy)

    HW, cpg = H * H, O // G
    m = cpg * HW
    zz = z.detach().view(N, G, cpg, HW)
    mean = zz.mean(dim=(2, 3))
    invstd = 1 / torch.sqrt((zz ** 2).mean(dim=(2, 3)) - mean ** 2 + eps)
    mu, inv = mean.repeat_interleave(cpg, 1), invstd.repeat_interleave(cpg, 1)      # [N, O]
    gam = gamma.detach()
    a = gam * inv
    b = beta.detach() - a * mu
    zf, dyf = z.detach().view(N, O, HW), dy.view(N, O, HW)
    pre = a[:, :, None] * zf + b[:, :, None]
    assert torch.allclose(torch.relu(pre), y.detach().view(N, O, HW), atol=1e-12)
    dym = dyf * (pre > 0)
    s1, s2 = dym.sum(2), (dym * zf).sum(2)
    t = inv * (s2 - mu * s1)
    A1 = ((gam * s1).view(N, G, cpg).sum(2) / m).repeat_interleave(cpg, 1)
    A2 = ((gam * t).view(N, G, cpg).sum(2) / m).repeat_interleave(cpg, 1)
    k1, k2, k3 = inv * gam, -inv * inv * A2, -inv * A1 + inv * inv * A2 * mu
    dz = k1[:, :, None] * dym + k2[:, :, None] * zf + k3[:, :, None]
    assert torch.allclose(dz, z.grad.view(N, O, HW), atol=1e-12)
    assert torch.allclose(t.sum(0), gamma.grad, atol=1e-12) and torch.allclose(s1.sum(0), beta.grad, atol=1e-12)


def test_batch_norm_backward_coefficients():
    """pointwise.cu bwd_coef_kernel: dz = k1 dy_m + k2 z + k3, dgamma = invstd (s2 - mean s1), dbeta = s1 against
    autograd of F.batch_norm (training statistics) + affine + ReLU."""
    torch.manual_seed(1)
    N, O, H, eps = 4, 8, 3, 1e-5
    z = torch.randn(N, O, H, H, dtype=torch.double, requires_grad=True)
    gamma = torch.randn(O, dtype=torch.double, requires_grad=True)
    beta = torch.randn(O, dtype=torch.double, requires_grad=True)
    y = torch.relu(gamma.view(1, -1, 1, 1) * F.batch_norm(z, None, None, None, None, True, 0.1, eps) + beta.view(1, -1, 1, 1))
    dy = torch.randn_like(y)
    y.backward(dy)
    zc = z.detach().permute(1, 0, 2, 3).reshape(O, -1)
    dyc = dy.permute(1, 0, 2, 3).reshape(O, -1)
    n = zc.shape[1]
    mean = zc.mean(1)
    invstd = 1 / torch.sqrt(zc.var(1, unbiased=False) + eps)
    gam = gamma.detach()
    a = gam * invstd
    b = beta.detach() - a * mean
    dym = dyc * ((a[:, None] * zc + b[:, None]) > 0)
    s1, s2 = dym.sum(1), (dym * zc).sum(1)
    dg = invstd * (s2 - mean * s1)
    c2 = -a * invstd * dg / n
    k1, k2, k3 = a, c2, -a * s1 / n - c2 * mean
    dz = k1[:, None] * dym + k2[:, None] * zc + k3[:, None]
    assert torch.allclose(dz, z.grad.permute(1, 0, 2, 3).reshape(O, -1), atol=1e-12)
    assert torch.allclose(dg, gamma.grad, atol=1e-12) and torch.allclose(s1, beta.grad, atol=1e-12)


def test_weight_gradient_tap_pairing_identity():
    """igemm_sm100.cu wgrad_om_kernel, paired mode (3x3 / stride 1 / pad 1): with dz shifted down by one image row
    (rows past the image are zero), the product against the x patches of tap (dh, dw) is the weight gradient of tap
    (dh - 1, dw):  sum_p dz[p + Q] x[p + (dh, dw)] = dW[(dh - 1, dw)]  — exact for dh - 1 = 0, where the dropped
    image row 0 only meets the zero padding.  So taps (1, .) yield dW[1, .] (lower rows) and dW[0, .] (upper rows)."""
    torch.manual_seed(2)
    N, C, O, H, W = 2, 3, 4, 6, 5
    x = torch.randn(N, C, H, W, dtype=torch.double)
    w = torch.randn(O, C, 3, 3, dtype=torch.double, requires_grad=True)
    dz = torch.randn(N, O, H, W, dtype=torch.double)
    F.conv2d(x, w, None, 1, 1).backward(dz)
    patches = F.unfold(x, 3, padding=1).view(N, C, 3, 3, H, W)          # x[p + (dh, dw) - pad]
    dz_shift = torch.zeros_like(dz)
    dz_shift[:, :, :-1] = dz[:, :, 1:]                                   # dz one image row further down, zero fill
    for dw_ in range(3):
        lower = torch.einsum('nohw,nchw->oc', dz, patches[:, :, 1, dw_])
        upper = torch.einsum('nohw,nchw->oc', dz_shift, patches[:, :, 1, dw_])
        assert torch.allclose(lower, w.grad[:, :, 1, dw_], atol=1e-12)
        assert torch.allclose(upper, w.grad[:, :, 0, dw_], atol=1e-12)
        # the same pairing one row lower (taps (2, .)) reproduces dW[1, .] only up to the dropped row 0 term, which is
        # why the kernel discards those upper rows
        upper2 = torch.einsum('nohw,nchw->oc', dz_shift, patches[:, :, 2, dw_])
        missing = torch.einsum('now,ncw->oc', dz[:, :, 0], patches[:, :, 1, dw_, 0])
        assert torch.allclose(upper2 + missing, w.grad[:, :, 1, dw_], atol=1e-12)


def test_strided_data_gradient_phase_decomposition():
    """api.cu plan_dgrad_phase: the data gradient of a stride-s conv, computed per output phase (h % s, w % s) with only
    the taps of that phase (no zero insertion), equals autograd's."""
    torch.manual_seed(3)
    N, C, O, H, k, s, p = 2, 3, 4, 9, 3, 2, 1
    x = torch.randn(N, C, H, H, dtype=torch.double, requires_grad=True)
    w = torch.randn(O, C, k, k, dtype=torch.double)
    z = F.conv2d(x, w, None, s, p)
    dz = torch.randn_like(z)
    z.backward(dz)
    P = z.shape[2]
    dx = torch.zeros(N, C, H, H, dtype=torch.double)
    for ph in range(s):
        for pw in range(s):
            for r in range(k):
                if (ph + p - r) % s:
                    continue
                for c in range(k):
                    if (pw + p - c) % s:
                        continue
                    for h in range(ph, H, s):
                        e = (h + p - r) // s
                        if not 0 <= e < P:
                            continue
                        for ww in range(pw, H, s):
                            f = (ww + p - c) // s
                            if 0 <= f < P:
                                dx[:, :, h, ww] += dz[:, :, e, f] @ w[:, :, r, c]
    assert torch.allclose(dx, x.grad, atol=1e-12)


def test_tf32_operand_model_of_the_oracle():
    """oracle.batch_conv(mode='tf32') — the model of BASELINE config 2 — cuts the operands of all three contractions of
    a convolution to TF32 and accumulates exactly as the fp32 operators do on those operands."""
    import torch.nn.functional as F
    from oracle import passport_oracle as po
    torch.manual_seed(0)
    x = torch.randn(3, 32, 6, 6, requires_grad=True)
    w = torch.randn(64, 32, 3, 3, requires_grad=True)
    b = torch.randn(64, requires_grad=True)
    g = torch.randn(3, 64, 6, 6)
    z = po.batch_conv(x, w, b, 1, 1, 'tf32')
    z.backward(g)
    xc, wc, gc = po.tf32_cut(x), po.tf32_cut(w), po.tf32_cut(g)
    # cut values are representable with 10 explicit mantissa bits and never larger in magnitude than the input
    assert torch.equal(po.tf32_cut(xc), xc) and bool((xc.abs() <= x.detach().abs()).all())
    assert float(((xc - x.detach()).abs() / x.detach().abs().clamp_min(1e-30)).max()) < 2.0 ** -10
    assert torch.equal(z.detach(), F.conv2d(xc, wc, b.detach(), 1, 1))
    assert torch.equal(x.grad, torch.nn.grad.conv2d_input(x.shape, wc, gc, 1, 1))
    assert torch.equal(w.grad, torch.nn.grad.conv2d_weight(xc, w.shape, gc, 1, 1))
    assert torch.allclose(b.grad, g.sum((0, 2, 3)))
    # and the whole thing stays within TF32 distance of the exact fp32 operator
    assert float((z.detach() - F.conv2d(x.detach(), w.detach(), b.detach(), 1, 1)).norm() / z.detach().norm()) < 2e-3

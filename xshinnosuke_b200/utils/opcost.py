"""Algorithmic cost (FLOPs / HBM bytes) of one C-ABI call, from its arguments -- the numerators of the
rooflines in bench.py (definitions: SURVEY.md 8d; e = 4 bytes, arg-max / masks 1 byte or 1 bit per element).

cost(name, args) -> (kind, amount) with kind "flop" (tensor-core bound entries) or "byte" (HBM bound)."""


def _conv_flops(a, off):
    # (..., N, H, W, C, OC, KH, KW, sh, sw, pad, ...) starting at args[off]
    N, H, W, C, OC, KH, KW, sh, sw, pad = a[off:off + 10]
    OH = (H - KH + 2 * pad) // sh + 1
    OW = (W - KW + 2 * pad) // sw + 1
    return 2.0 * N * OH * OW * C * KH * KW * OC


def _pool_out(H, W, k, sh, sw, pad):
    return (H + 2 * pad - k) // sh + 1, (W + 2 * pad - k) // sw + 1


def cost(name, a):
    if name in ("xsb_conv2d_fwd", "xsb_conv2d_fwd_stats", "xsb_conv2d_fwd_acc"):
        return "flop", _conv_flops(a, 4)
    if name == "xsb_split_tf32":
        return "byte", 12.0 * a[3]
    if name in ("xsb_conv2d_dgrad", "xsb_conv2d_wgrad"):
        return "flop", _conv_flops(a, 3)
    if name == "xsb_gemm":
        return "flop", 2.0 * a[4] * a[5] * a[6]
    if name == "xsb_relu_fwd":
        n = a[3]
        return "byte", 8.0 * n + (n / 8.0 if a[2] else 0.0)
    if name == "xsb_relu_bwd":
        return "byte", 12.0 * a[3]
    if name == "xsb_relu_bwd_mask":
        return "byte", 8.0 * a[3] + a[3] / 8.0
    if name == "xsb_add":
        return "byte", 12.0 * a[3]
    if name == "xsb_axpy":
        return "byte", 12.0 * a[3]
    if name == "xsb_fill":
        return "byte", 4.0 * a[2]
    if name in ("xsb_nchw_to_nhwc", "xsb_nhwc_to_nchw"):
        return "byte", 8.0 * a[2] * a[3] * a[4] * a[5]
    if name == "xsb_maxpool2d_fwd":
        N, H, W, C, k, sh, sw, pad = a[3:11]
        OH, OW = _pool_out(H, W, k, sh, sw, pad)
        return "byte", 4.0 * (N * H * W * C + N * OH * OW * C) + (N * OH * OW * C if a[2] else 0)
    if name == "xsb_bn_relu_maxpool_fwd":
        N, H, W, C, k, sh, sw, pad = a[8:16]
        OH, OW = _pool_out(H, W, k, sh, sw, pad)
        return "byte", 4.0 * (N * H * W * C + N * OH * OW * C) + (N * OH * OW * C if a[6] else 0) + \
            (N * H * W * C / 8.0 if a[7] else 0.0)
    if name == "xsb_maxpool2d_bwd":
        N, H, W, C, k, sh, sw, pad = a[3:11]
        OH, OW = _pool_out(H, W, k, sh, sw, pad)
        return "byte", 4.0 * (N * H * W * C + N * OH * OW * C) + N * OH * OW * C
    if name == "xsb_avgpool2d_fwd":
        N, H, W, C, k, sh, sw = a[2:9]
        OH, OW = _pool_out(H, W, k, sh, sw, 0)
        return "byte", 4.0 * (N * H * W * C + N * OH * OW * C)
    if name == "xsb_avgpool2d_bwd":
        N, H, W, C, k, sh, sw = a[2:9]
        OH, OW = _pool_out(H, W, k, sh, sw, 0)
        return "byte", 4.0 * (N * H * W * C + N * OH * OW * C)
    if name == "xsb_bn_train_fwd":
        return "byte", 12.0 * a[8] * a[9]
    if name == "xsb_bn_stats":
        return "byte", 4.0 * a[5] * a[6]
    if name == "xsb_bn_apply":
        return "byte", 8.0 * a[6] * a[7] + (a[6] * a[7] / 8.0 if a[8] else 0.0)
    if name == "xsb_bn_add_relu":          # BatchNorm input + shortcut read, rectified sum + bit mask written
        return "byte", 12.0 * a[12] * a[13] + a[12] * a[13] / 8.0
    if name == "xsb_add_relu":
        return "byte", 12.0 * a[4] + a[4] / 8.0
    if name in ("xsb_bn_bwd", "xsb_bn_bwd_dxsum"):
        return "byte", (20.0 if a[5] else 8.0) * a[8] * a[9]
    if name == "xsb_bn_bwd_dxsum_side":     # + the masked dy written out for the identity shortcut
        return "byte", 24.0 * a[8] * a[9]
    if name == "xsb_bn_bwd_pool_dxsum":
        # two passes over (x, dpool, argmax, mask) + dx; the scattered dy the unfused pair writes and reads twice is gone
        N, H, W, C = a[9:13]
        OH, OW = _pool_out(H, W, 3, 2, 2, 1)
        n, m = float(N) * H * W * C, float(N) * OH * OW * C
        return "byte", 2.0 * (4.0 * n + 5.0 * m + (n / 8.0 if a[16] else 0.0)) + 4.0 * n
    if name == "xsb_bn_finalize":
        return "byte", 12.0 * a[1] * a[3]      # P partials x {n, mean, M2} x C
    if name == "xsb_colsum":
        return "byte", 4.0 * a[2] * a[3]
    if name == "xsb_sgd_step":
        return "byte", 12.0 * a[2]
    if name == "xsb_adam_step":
        return "byte", 28.0 * a[4]
    if name in ("xsb_softmax_xent_fwd",):
        return "byte", 8.0 * a[4] * a[5]
    if name in ("xsb_softmax_xent_bwd",):
        return "byte", 8.0 * a[4] * a[5]
    if name in ("xsb_reduce_axis",):
        return "byte", 4.0 * a[2] * a[3] * a[4]
    if name in ("xsb_broadcast_axis",):
        return "byte", 4.0 * a[2] * a[3] * a[4]
    return "byte", 0.0

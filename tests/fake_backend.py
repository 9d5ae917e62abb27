"""TEST INFRASTRUCTURE ONLY: a NumPy stand-in for libxsb200.so so that the HOST logic of
xshinnosuke_b200 (graph wiring, autograd walk, lazy-zero gradients, flat optimizer buffers, static
graph, data-parallel buckets) can be exercised by `-m "not gpu"` tests on a box without a GPU.

Every function takes the same arguments as its C-ABI namesake (include/xsb200.h), with "device"
pointers being addresses of CPU torch tensors, and computes the kernel's contract in float64
(using the oracle where one exists). The product package never imports this module; GPU tests never
install it. `install()` monkeypatches `xshinnosuke_b200._lib.lib` and the runtime singleton.
"""
import ctypes

import numpy as np
import torch

from oracle import xs_numpy as X


def _view(ptr, n, ctype, dtype):
    if ptr is None or n == 0:
        return None
    return np.ctypeslib.as_array((ctype * int(n)).from_address(int(ptr))).view(dtype)


def f32(ptr, n):
    return _view(ptr, n, ctypes.c_float, np.float32)


def u8(ptr, n):
    return _view(ptr, n, ctypes.c_uint8, np.uint8)


def i64(ptr, n):
    return _view(ptr, n, ctypes.c_int64, np.int64)


def i32(ptr, n):
    return _view(ptr, n, ctypes.c_int32, np.int32)


def _store(dst, val, acc):
    val = np.asarray(val, dtype=np.float64).reshape(dst.shape)
    dst[...] = (dst.astype(np.float64) + val if acc else val).astype(np.float32)


class FakeLib:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if not name.startswith("xsb_"):
            raise AttributeError(name)
        impl = getattr(self, "_" + name[4:])

        def wrapped(*a):
            self.calls.append(name)
            r = impl(*a)
            return 0 if r is None else r
        return wrapped

    # ---- runtime
    def _version(self):
        return 100

    def _init(self, dev):
        pass

    def _device_sm_count(self):
        return 148

    def _stream_sync(self, s):
        pass

    # ---- elementwise
    def _relu_fwd(self, x, y, mask, n, s):
        xv = f32(x, n).copy()
        f32(y, n)[...] = np.maximum(xv, 0)
        if mask:
            u8(mask, (n + 7) // 8)[...] = np.packbits(xv < 0, bitorder="little")

    def _relu_bwd(self, dy, x, dx, n, acc, s):
        g = f32(dy, n).astype(np.float64)
        g[f32(x, n) < 0] = 0
        _store(f32(dx, n), g, acc)

    def _relu_bwd_mask(self, dy, mask, dx, n, acc, s):
        m = np.unpackbits(u8(mask, (n + 7) // 8), bitorder="little")[:n].astype(bool)
        g = f32(dy, n).astype(np.float64)
        g[m] = 0
        _store(f32(dx, n), g, acc)

    def _add(self, a, b, y, n, s):
        f32(y, n)[...] = f32(a, n) + f32(b, n)

    def _axpy(self, dst, src, alpha, n, s):
        f32(dst, n)[...] += np.float32(alpha) * f32(src, n)

    def _fill(self, dst, v, n, s):
        f32(dst, n)[...] = v

    def _nchw_to_nhwc(self, src, dst, N, C, H, W, acc, s):
        a = f32(src, N * C * H * W).reshape(N, C, H, W).transpose(0, 2, 3, 1)
        _store(f32(dst, N * C * H * W).reshape(N, H, W, C), a, acc)

    def _nhwc_to_nchw(self, src, dst, N, C, H, W, acc, s):
        a = f32(src, N * C * H * W).reshape(N, H, W, C).transpose(0, 3, 1, 2)
        _store(f32(dst, N * C * H * W).reshape(N, C, H, W), a, acc)

    def _reduce_axis(self, x, y, outer, red, inner, scale, s):
        a = f32(x, outer * red * inner).reshape(outer, red, inner).astype(np.float64)
        f32(y, outer * inner)[...] = (a.sum(axis=1) * scale).reshape(-1)

    def _broadcast_axis(self, dy, dx, outer, red, inner, scale, acc, s):
        g = f32(dy, outer * inner).reshape(outer, 1, inner).astype(np.float64) * scale
        _store(f32(dx, outer * red * inner).reshape(outer, red, inner), np.broadcast_to(g, (outer, red, inner)), acc)

    # ---- pooling (channels-last buffers <-> oracle NCHW)
    @staticmethod
    def _nchw(ptr, N, H, W, C):
        return f32(ptr, N * H * W * C).reshape(N, H, W, C).transpose(0, 3, 1, 2)

    def _maxpool2d_fwd(self, x, y, amax, N, H, W, C, k, sh, sw, pad, s):
        yv, am = X.max_pool2d_fwd(self._nchw(x, N, H, W, C), k, (sh, sw), pad)
        OH, OW = yv.shape[2:]
        f32(y, N * OH * OW * C)[...] = yv.transpose(0, 2, 3, 1).reshape(-1)
        if amax:
            u8(amax, am.size)[...] = am.astype(np.uint8)

    def _maxpool2d_bwd(self, dy, amax, dx, N, H, W, C, k, sh, sw, pad, acc, s):
        OH, OW = (H + 2 * pad - k) // sh + 1, (W + 2 * pad - k) // sw + 1
        g = self._nchw(dy, N, OH, OW, C)
        am = u8(amax, N * OH * OW * C).astype(np.int64)
        d = X.max_pool2d_bwd(g, am, (N, C, H, W), k, (sh, sw), pad)
        _store(f32(dx, N * H * W * C).reshape(N, H, W, C), d.transpose(0, 2, 3, 1), acc)

    def _avgpool2d_fwd(self, x, y, N, H, W, C, k, sh, sw, s):
        yv = X.avg_pool2d_fwd(self._nchw(x, N, H, W, C), k, (sh, sw))
        f32(y, yv.size)[...] = yv.transpose(0, 2, 3, 1).reshape(-1)

    def _avgpool2d_bwd(self, dy, dx, N, H, W, C, k, sh, sw, acc, s):
        OH, OW = (H - k) // sh + 1, (W - k) // sw + 1
        d = X.avg_pool2d_bwd(self._nchw(dy, N, OH, OW, C), (N, C, H, W), k, (sh, sw))
        _store(f32(dx, N * H * W * C).reshape(N, H, W, C), d.transpose(0, 2, 3, 1), acc)

    # ---- batch norm on [R, C]
    def _bn_workspace_floats(self, R, C):
        return 64

    def _bn_train_fwd(self, x, gamma, beta, rm, rv, y, smean, sinv, R, C, eps, mom, ws, s):
        xv = f32(x, R * C).reshape(R, C).astype(np.float64)
        mean, var = xv.mean(axis=0), xv.var(axis=0)
        inv = 1.0 / np.sqrt(var + eps)
        f32(y, R * C).reshape(R, C)[...] = (xv - mean) * inv * f32(gamma, C) + f32(beta, C)
        f32(smean, C)[...] = mean
        f32(sinv, C)[...] = inv
        if rm:
            f32(rm, C)[...] = mom * f32(rm, C).astype(np.float64) + (1 - mom) * mean
        if rv:
            f32(rv, C)[...] = mom * f32(rv, C).astype(np.float64) + (1 - mom) * var

    def _bn_stats(self, x, rm, rv, smean, sinv, R, C, eps, mom, ws, s):
        xv = f32(x, R * C).reshape(R, C).astype(np.float64)
        mean, var = xv.mean(axis=0), xv.var(axis=0)
        f32(smean, C)[...] = mean
        f32(sinv, C)[...] = 1.0 / np.sqrt(var + eps)
        if rm:
            f32(rm, C)[...] = mom * f32(rm, C).astype(np.float64) + (1 - mom) * mean
        if rv:
            f32(rv, C)[...] = mom * f32(rv, C).astype(np.float64) + (1 - mom) * var

    def _bn_apply(self, x, y, mean, invstd, gamma, beta, R, C, relu, mask, s):
        xv = f32(x, R * C).reshape(R, C).astype(np.float64)
        o = ((xv - f32(mean, C)) * f32(invstd, C) * f32(gamma, C) + f32(beta, C)).astype(np.float32)
        if relu:
            u8(mask, (R * C + 7) // 8)[...] = np.packbits(o.reshape(-1) < 0, bitorder="little")
            o = np.maximum(o, 0)
        f32(y, R * C).reshape(R, C)[...] = o

    def _bn_relu_maxpool_fwd(self, x, mean, invstd, gamma, beta, y, amax, rmask, N, H, W, C, k, sh, sw, pad, s):
        # contract: maxpool(max(0, bn(x))) with argmax + the ReLU bit mask, the intermediate is never stored
        R = N * H * W
        z = torch.empty(R * C, dtype=torch.float32)
        m = torch.empty((R * C + 7) // 8, dtype=torch.uint8)
        self._bn_apply(x, z.data_ptr(), mean, invstd, gamma, beta, R, C, 1, m.data_ptr(), s)
        if rmask:
            u8(rmask, (R * C + 7) // 8)[...] = m.numpy()
        self._maxpool2d_fwd(z.data_ptr(), y, amax, N, H, W, C, k, sh, sw, pad, s)
        return 0

    def _bn_add_relu(self, xa, mean_a, invstd_a, gamma_a, beta_a, b, mean_b, invstd_b, gamma_b, beta_b, z, mask, R, C, s):
        # contract: xsb_bn_apply(xa) [+ xsb_bn_apply(b)] into scratch, then xsb_add_relu
        n = R * C
        ta = torch.empty(n, dtype=torch.float32)
        self._bn_apply(xa, ta.data_ptr(), mean_a, invstd_a, gamma_a, beta_a, R, C, 0, None, s)
        rhs = b
        if mean_b:
            tb = torch.empty(n, dtype=torch.float32)
            self._bn_apply(b, tb.data_ptr(), mean_b, invstd_b, gamma_b, beta_b, R, C, 0, None, s)
            rhs = tb.data_ptr()
        self._add_relu(ta.data_ptr(), rhs, z, mask, n, s)
        return 0

    def _add_relu(self, a, b, y, mask, n, s):
        t = f32(a, n) + f32(b, n)
        u8(mask, (n + 7) // 8)[...] = np.packbits(t < 0, bitorder="little")
        f32(y, n)[...] = np.maximum(t, 0)

    def _bn_bwd(self, dy, x, gamma, smean, sinv, dx, dgamma, dbeta, R, C, acc_dx, acc_dg, acc_db, rmask, ws, s):
        g = f32(dy, R * C).reshape(R, C).astype(np.float64)
        if rmask:
            m = np.unpackbits(u8(rmask, (R * C + 7) // 8), bitorder="little")[:R * C].astype(bool).reshape(R, C)
            g = np.where(m, 0.0, g)
        xhat = (f32(x, R * C).reshape(R, C).astype(np.float64) - f32(smean, C)) * f32(sinv, C)
        sdy, sdyx = g.sum(axis=0), (g * xhat).sum(axis=0)
        if dgamma:
            _store(f32(dgamma, C), sdyx, acc_dg)
        if dbeta:
            _store(f32(dbeta, C), sdy, acc_db)
        if dx:
            d = f32(gamma, C).astype(np.float64) * f32(sinv, C) * (g - sdy / R - xhat * sdyx / R)
            _store(f32(dx, R * C).reshape(R, C), d, acc_dx)

    def _bn_bwd_dxsum(self, dy, x, gamma, smean, sinv, dx, dgamma, dbeta, R, C, acc_dx, acc_dg, acc_db, rmask, ws,
                      dxsum, acc_dxsum, done, s):
        # contract: dxsum (+)= column sums of the dx CONTRIBUTION; done (host int) = 1
        before = f32(dx, R * C).reshape(R, C).astype(np.float64).copy() if acc_dx else 0.0
        self._bn_bwd(dy, x, gamma, smean, sinv, dx, dgamma, dbeta, R, C, acc_dx, acc_dg, acc_db, rmask, ws, s)
        contrib = f32(dx, R * C).reshape(R, C).astype(np.float64) - before
        _store(f32(dxsum, C), contrib.sum(axis=0), acc_dxsum)
        done._obj.value = 1

    def _bn_bwd_dxsum_side(self, dy, x, gamma, smean, sinv, dx, dgamma, dbeta, R, C, acc_dx, acc_dg, acc_db, rmask, ws,
                           dxsum, acc_dxsum, done, side, side_done, s):
        # contract: xsb_bn_bwd_dxsum (dxsum nullable) + side = dy with the masked elements zeroed; side_done (host int) = 1
        if dxsum:
            self._bn_bwd_dxsum(dy, x, gamma, smean, sinv, dx, dgamma, dbeta, R, C, acc_dx, acc_dg, acc_db, rmask, ws, dxsum,
                               acc_dxsum, done, s)
        else:
            self._bn_bwd(dy, x, gamma, smean, sinv, dx, dgamma, dbeta, R, C, acc_dx, acc_dg, acc_db, rmask, ws, s)
        g = f32(dy, R * C).astype(np.float64)
        if rmask:
            m = np.unpackbits(u8(rmask, (R * C + 7) // 8), bitorder="little")[:R * C].astype(bool)
            g = np.where(m, 0.0, g)
        f32(side, R * C)[...] = g
        side_done._obj.value = 1

    def _bn_bwd_pool_dxsum(self, dpool, amax, x, gamma, smean, sinv, dx, dgamma, dbeta, N, H, W, C, acc_dx, acc_dg, acc_db,
                           rmask, ws, dxsum, acc_dxsum, s):
        # contract: xsb_maxpool2d_bwd (3x3, stride 2, pad 1) into a scratch dy, then xsb_bn_bwd(_dxsum) on it
        R = N * H * W
        dy = torch.zeros(R * C, dtype=torch.float32)
        self._maxpool2d_bwd(dpool, amax, dy.data_ptr(), N, H, W, C, 3, 2, 2, 1, 0, s)
        before = f32(dx, R * C).reshape(R, C).astype(np.float64).copy() if acc_dx else 0.0
        self._bn_bwd(dy.data_ptr(), x, gamma, smean, sinv, dx, dgamma, dbeta, R, C, acc_dx, acc_dg, acc_db, rmask, ws, s)
        if dxsum:
            contrib = f32(dx, R * C).reshape(R, C).astype(np.float64) - before
            _store(f32(dxsum, C), contrib.sum(axis=0), acc_dxsum)
        return 0

    def _bn_finalize(self, part, P, R, C, rm, rv, smean, sinv, eps, mom, s):
        pv = f32(part, P * 3 * C).reshape(P, 3, C).astype(np.float64)
        n, mean, m2 = pv[:, 0], pv[:, 1], pv[:, 2]
        tot = n.sum(axis=0)
        gm = (n * mean).sum(axis=0) / tot
        var = (m2 + n * (mean - gm) ** 2).sum(axis=0) / R
        assert np.all(tot == R)
        f32(smean, C)[...] = gm
        f32(sinv, C)[...] = 1.0 / np.sqrt(var + eps)
        if rm:
            f32(rm, C)[...] = mom * f32(rm, C).astype(np.float64) + (1 - mom) * gm
        if rv:
            f32(rv, C)[...] = mom * f32(rv, C).astype(np.float64) + (1 - mom) * var

    def _colsum(self, x, out, R, C, acc, ws, s):
        _store(f32(out, C), f32(x, R * C).reshape(R, C).astype(np.float64).sum(axis=0), acc)

    # ---- GEMM / conv
    def _gemm(self, A, B, C, bias, M, N, K, ta, tb, acc, mode, ws, wsn, s):
        a = f32(A, M * K).reshape((K, M) if ta else (M, K)).astype(np.float64)
        b = f32(B, K * N).reshape((N, K) if tb else (K, N)).astype(np.float64)
        r = (a.T if ta else a) @ (b.T if tb else b)
        if bias:
            r = r + f32(bias, N)
        _store(f32(C, M * N).reshape(M, N), r, acc)

    @staticmethod
    def _w_oihw(ptr, OC, KH, KW, C):
        return f32(ptr, OC * KH * KW * C).reshape(OC, KH, KW, C).transpose(0, 3, 1, 2)

    def _conv2d_fwd(self, x, w, bias, y, N, H, W, C, OC, KH, KW, sh, sw, pad, mode, ws, wsn, s):
        b = f32(bias, OC).reshape(1, OC).astype(np.float64) if bias else None
        yv, _ = X.conv2d_fwd(self._nchw(x, N, H, W, C).astype(np.float64), self._w_oihw(w, OC, KH, KW, C).astype(np.float64),
                             b, (sh, sw), pad)
        f32(y, yv.size)[...] = yv.transpose(0, 2, 3, 1).reshape(-1)

    def _conv2d_fwd_relu(self, x, w, bias, y, act, N, H, W, C, OC, KH, KW, sh, sw, pad, mode, ws, wsn, s):
        self._conv2d_fwd(x, w, bias, y, N, H, W, C, OC, KH, KW, sh, sw, pad, mode, ws, wsn, s)
        OH, OW = (H - KH + 2 * pad) // sh + 1, (W - KW + 2 * pad) // sw + 1
        f32(act, N * OH * OW * OC)[...] = np.maximum(f32(y, N * OH * OW * OC), 0.0)

    def _gemm_bias_relu(self, A, B, C, bias, act, M, N, K, ta, tb, mode, ws, wsn, s):
        self._gemm(A, B, C, bias, M, N, K, ta, tb, 0, mode, ws, wsn, s)
        f32(act, M * N)[...] = np.maximum(f32(C, M * N), 0.0)

    def _conv2d_fwd_acc(self, x, w, bias, y, N, H, W, C, OC, KH, KW, sh, sw, pad, acc, mode, ws, wsn, s):
        OH, OW = (H - KH + 2 * pad) // sh + 1, (W - KW + 2 * pad) // sw + 1
        old = f32(y, N * OH * OW * OC).astype(np.float64).copy() if acc else 0.0
        self._conv2d_fwd(x, w, bias, y, N, H, W, C, OC, KH, KW, sh, sw, pad, mode, ws, wsn, s)
        if acc:
            f32(y, N * OH * OW * OC)[...] = f32(y, N * OH * OW * OC).astype(np.float64) + old

    def _split_tf32(self, x, hi, lo, n, s):
        def rnd(a):      # round to tf32 (10 mantissa bits), nearest
            u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
            return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)
        v = f32(x, n)
        h = rnd(v)
        f32(hi, n)[...] = h
        f32(lo, n)[...] = rnd(v - h)

    def _conv2d_fwd_stats(self, x, w, bias, y, N, H, W, C, OC, KH, KW, sh, sw, pad, mode, ws, wsn, stats, stats_n,
                          n_partials, s):
        # contract: y as xsb_conv2d_fwd; stats[p][{n, mean, M2}][OC] partials (two here, split over the rows)
        self._conv2d_fwd(x, w, bias, y, N, H, W, C, OC, KH, KW, sh, sw, pad, mode, ws, wsn, s)
        OH, OW = (H - KH + 2 * pad) // sh + 1, (W - KW + 2 * pad) // sw + 1
        R = N * OH * OW
        yv = f32(y, R * OC).reshape(R, OC).astype(np.float64)
        halves = [yv[: R // 2], yv[R // 2:]] if R > 1 else [yv]
        if stats_n < len(halves) * 3 * OC:
            n_partials._obj.value = 0
            return
        pv = f32(stats, len(halves) * 3 * OC).reshape(len(halves), 3, OC)
        for i, hv in enumerate(halves):
            pv[i, 0] = hv.shape[0]
            pv[i, 1] = hv.mean(axis=0)
            pv[i, 2] = ((hv - hv.mean(axis=0)) ** 2).sum(axis=0)
        n_partials._obj.value = len(halves)

    def _conv_bwd(self, dy, x, w, N, H, W, C, OC, KH, KW, sh, sw, pad, need_dx):
        OH, OW = (H - KH + 2 * pad) // sh + 1, (W - KW + 2 * pad) // sw + 1
        g = self._nchw(dy, N, OH, OW, OC).astype(np.float64)
        xv = self._nchw(x, N, H, W, C).astype(np.float64) if x else np.zeros((N, C, H, W))
        wv = self._w_oihw(w, OC, KH, KW, C).astype(np.float64) if w else np.zeros((OC, C, KH, KW))
        col = X.unfold(np.pad(xv, ((0, 0), (0, 0), (pad, pad), (pad, pad))), OH, OW, KH, KW, (sh, sw))
        return X.conv2d_bwd(g, (N, C, H, W), wv, col, (sh, sw), pad, need_dx)

    def _conv2d_dgrad(self, dy, w, dx, N, H, W, C, OC, KH, KW, sh, sw, pad, acc, mode, ws, wsn, s):
        d, _, _ = self._conv_bwd(dy, None, w, N, H, W, C, OC, KH, KW, sh, sw, pad, True)
        _store(f32(dx, N * H * W * C).reshape(N, H, W, C), d.transpose(0, 2, 3, 1), acc)

    def _conv2d_wgrad(self, dy, x, dw, N, H, W, C, OC, KH, KW, sh, sw, pad, acc, mode, ws, wsn, s):
        _, g, _ = self._conv_bwd(dy, x, None, N, H, W, C, OC, KH, KW, sh, sw, pad, False)
        _store(f32(dw, OC * KH * KW * C).reshape(OC, KH, KW, C), g.transpose(0, 2, 3, 1), acc)

    # ---- losses
    def _softmax_xent_fwd(self, z, labels, probs, loss, N, C, red, bad, s):
        zv = f32(z, N * C).reshape(N, C).astype(np.float64)
        lab = i64(labels, N)
        if ((lab < 0) | (lab >= C)).any():
            f32(loss, 1)[0] = np.nan
            if bad:
                i32(bad, 1)[0] = 1
            return
        l, p = X.cross_entropy_fwd(zv, lab, "mean" if red == 0 else "sum")
        f32(probs, N * C)[...] = p.reshape(-1)
        f32(loss, 1)[0] = l[0]

    def _softmax_xent_bwd(self, probs, labels, dloss, dz, N, C, red, acc, s):
        p = f32(probs, N * C).reshape(N, C).astype(np.float64)
        g = X.cross_entropy_bwd(p, i64(labels, N), float(f32(dloss, 1)[0]), "mean" if red == 0 else "sum")
        _store(f32(dz, N * C).reshape(N, C), g, acc)

    def _mse_fwd(self, a, b, loss, n, red, s):
        f32(loss, 1)[0] = X.mse_loss_fwd(f32(a, n).astype(np.float64), f32(b, n).astype(np.float64),
                                         "mean" if red == 0 else "sum")[0]

    def _mse_bwd(self, a, b, dloss, da, db, n, red, acc_a, acc_b, s):
        g = X.mse_loss_bwd(f32(a, n).astype(np.float64), f32(b, n).astype(np.float64), float(f32(dloss, 1)[0]),
                           "mean" if red == 0 else "sum")
        if da:
            _store(f32(da, n), g, acc_a)
        if db:
            _store(f32(db, n), -g, acc_b)

    # ---- SURVEY 8f-4 ops
    def _unary_fwd(self, kind, x, y, n, s):
        v = f32(x, n).astype(np.float64)
        f32(y, n)[...] = X.sigmoid_fwd(v) if kind == 0 else X.tanh_fwd(v)

    def _unary_bwd(self, kind, dy, y, dx, n, acc, s):
        g, o = f32(dy, n).astype(np.float64), f32(y, n).astype(np.float64)
        _store(f32(dx, n), X.sigmoid_bwd(g, o) if kind == 0 else X.tanh_bwd(g, o), acc)

    def _dropout_fwd(self, x, y, mask, n, keep, seed, step, s):
        rs = np.random.RandomState(int(seed) % (2 ** 31))
        m = (rs.rand(n) < keep)
        bits = np.packbits(np.concatenate([m, np.zeros((-n) % 8, dtype=bool)]), bitorder="little")
        u8(mask, len(bits))[...] = bits
        f32(y, n)[...] = f32(x, n) * m / keep

    def _dropout_bwd(self, dy, mask, dx, n, scale, acc, s):
        m = np.unpackbits(u8(mask, (n + 7) // 8), bitorder="little")[:n]
        _store(f32(dx, n), f32(dy, n).astype(np.float64) * m * scale, acc)

    def _pad2d_fwd(self, x, y, N, H, W, C, ph, pw, s):
        v = f32(x, N * H * W * C).reshape(N, H, W, C)
        f32(y, N * (H + 2 * ph) * (W + 2 * pw) * C).reshape(N, H + 2 * ph, W + 2 * pw, C)[...] = \
            np.pad(v, ((0, 0), (ph, ph), (pw, pw), (0, 0)), "constant")

    def _pad2d_bwd(self, dy, dx, N, H, W, C, ph, pw, acc, s):
        g = f32(dy, N * (H + 2 * ph) * (W + 2 * pw) * C).reshape(N, H + 2 * ph, W + 2 * pw, C)
        _store(f32(dx, N * H * W * C).reshape(N, H, W, C), g[:, ph:ph + H, pw:pw + W, :].astype(np.float64), acc)

    def _layernorm_fwd(self, x, gamma, beta, y, mean, rstd, N, F, eps, s):
        v = f32(x, N * F).reshape(N, F).astype(np.float64)
        out, xhat, std = X.layer_norm_fwd(v, f32(gamma, F).astype(np.float64), f32(beta, F).astype(np.float64), eps)
        f32(y, N * F)[...] = out.reshape(-1)
        f32(mean, N)[...] = v.mean(axis=1)
        f32(rstd, N)[...] = 1.0 / std.reshape(-1)

    def _layernorm_bwd(self, dy, x, gamma, mean, rstd, dx, dgamma, dbeta, N, F, acc_dx, acc_g, acc_b, s):
        v = f32(x, N * F).reshape(N, F).astype(np.float64)
        g = f32(dy, N * F).reshape(N, F).astype(np.float64)
        mu, rs = f32(mean, N).astype(np.float64).reshape(N, 1), f32(rstd, N).astype(np.float64).reshape(N, 1)
        xhat = (v - mu) * rs
        d, dg, db = X.layer_norm_bwd(g, xhat, 1.0 / rs, f32(gamma, F).astype(np.float64))
        if dx:
            _store(f32(dx, N * F).reshape(N, F), d, acc_dx)
        if dgamma:
            _store(f32(dgamma, F), dg, acc_g)
        if dbeta:
            _store(f32(dbeta, F), db, acc_b)

    def _groupnorm_fwd(self, x, gamma, beta, y, mean, rstd, N, HW, C, G, eps, s):
        v = f32(x, N * HW * C).reshape(N, HW, 1, C).transpose(0, 3, 1, 2).astype(np.float64)      # [N, C, HW, 1]
        out, xhat, std = X.group_norm_fwd(v, f32(gamma, C).astype(np.float64), f32(beta, C).astype(np.float64), eps, G)
        f32(y, N * HW * C).reshape(N, HW, 1, C)[...] = out.transpose(0, 2, 3, 1)
        f32(mean, N * G)[...] = v.reshape(N, G, -1).mean(axis=2).reshape(-1)
        f32(rstd, N * G)[...] = 1.0 / std.reshape(-1)

    def _groupnorm_bwd(self, dy, x, gamma, mean, rstd, dx, dgamma, dbeta, N, HW, C, G, acc_dx, acc_g, acc_b, s):
        v = f32(x, N * HW * C).reshape(N, HW, 1, C).transpose(0, 3, 1, 2).astype(np.float64)
        g = f32(dy, N * HW * C).reshape(N, HW, 1, C).transpose(0, 3, 1, 2).astype(np.float64)
        mu = f32(mean, N * G).astype(np.float64).reshape(N, G, 1, 1, 1)
        rs = f32(rstd, N * G).astype(np.float64).reshape(N, G, 1, 1, 1)
        xhat = ((v.reshape(N, G, C // G, HW, 1) - mu) * rs).reshape(N, C, HW, 1)
        d, dg, db = X.group_norm_bwd(g, xhat, 1.0 / rs, f32(gamma, C).astype(np.float64), G)
        if dx:
            _store(f32(dx, N * HW * C).reshape(N, HW, 1, C), d.transpose(0, 2, 3, 1), acc_dx)
        if dgamma:
            _store(f32(dgamma, C), dg, acc_g)
        if dbeta:
            _store(f32(dbeta, C), db, acc_b)

    # ---- optimizers
    def _sgd_step(self, p, g, n, lr, gscale, wd, s):
        pv = f32(p, n).astype(np.float64)
        f32(p, n)[...] = pv - lr * (gscale * f32(g, n).astype(np.float64) + wd * pv)

    def _rule_step(self, kind, p, g, s1, s2, n, lr, rho, eps, gscale, wd, s):
        P = [f32(p, n).astype(np.float64)]
        G = [f32(g, n).astype(np.float64) * gscale + wd * P[0]]
        st = {"v": [f32(s1, n).astype(np.float64)], "s": [f32(s1, n).astype(np.float64)]}
        if kind == 3:
            st["x"] = [f32(s2, n).astype(np.float64)]
        if kind == 0:
            X.momentum_step(P, G, st, lr, rho)
        elif kind == 1:
            X.rmsprop_step(P, G, st, lr, rho, eps)
        elif kind == 2:
            X.adagrad_step(P, G, st, lr, eps)
        else:
            X.adadelta_step(P, G, st, rho, eps)
        f32(p, n)[...] = P[0]
        f32(s1, n)[...] = st["v"][0] if kind == 0 else st["s"][0]
        if kind == 3:
            f32(s2, n)[...] = st["x"][0]

    def _bn_eval_fwd(self, x, gamma, beta, mm, mv, y, R, C, eps, ws, s):
        xv = f32(x, R * C).reshape(R, C).astype(np.float64)
        f32(y, R * C).reshape(R, C)[...] = (xv - f32(mm, C)) / np.sqrt(f32(mv, C).astype(np.float64) + eps) * f32(gamma, C) \
            + f32(beta, C)

    def _tc_stats(self, out):
        pass

    def _counter_inc(self, c, s):
        i32(c, 1)[0] += 1

    def _adam_step(self, p, g, v, m, n, lr, b1, b2, eps, gscale, wd, step, s):
        t = int(i32(step, 1)[0])
        gv = f32(g, n).astype(np.float64) * gscale + wd * f32(p, n).astype(np.float64)
        vv = b1 * f32(v, n).astype(np.float64) + (1 - b1) * gv
        mv = b2 * f32(m, n).astype(np.float64) + (1 - b2) * gv * gv
        f32(p, n)[...] = f32(p, n).astype(np.float64) - lr * (vv / (1 - b1 ** t)) / (np.sqrt(mv / (1 - b2 ** t)) + eps)
        f32(v, n)[...] = vv
        f32(m, n)[...] = mv


def install(monkeypatch=None):
    """Point xshinnosuke_b200 at the fake library and a CPU 'device'. Returns the FakeLib."""
    import xshinnosuke_b200._lib as L
    import xshinnosuke_b200.core.device as D
    fake = FakeLib()
    targets = [L, D]
    import xshinnosuke_b200.nn.functional as F
    import xshinnosuke_b200.nn.grad_fn as G
    import xshinnosuke_b200.optim.optimizer as O
    import xshinnosuke_b200.layers.graph_ops as GO
    import xshinnosuke_b200.nn.x3 as X3
    targets += [F, G, O, GO, X3]
    try:
        import xshinnosuke_b200.parallel as P
        targets.append(P)
    except ImportError:
        pass
    saved = [(m, m.lib) for m in targets if hasattr(m, "lib")]
    for m, _ in saved:
        if monkeypatch is not None:
            monkeypatch.setattr(m, "lib", fake)
        else:
            m.lib = fake
    rt = D.rt
    state = dict(ready=rt.ready, device=rt.device, ws_small=rt._ws_small, ws_big=rt._ws_big, ws_stats=rt._ws_stats)

    def set_(obj, name, val):
        if monkeypatch is not None:
            monkeypatch.setattr(obj, name, val, raising=False)
        else:
            setattr(obj, name, val)
    set_(rt, "ready", True)
    set_(rt, "device", torch.device("cpu"))
    set_(rt, "_ws_small", torch.zeros(1024))
    set_(rt, "_ws_big", torch.zeros(1024))
    set_(rt, "_ws_stats", torch.zeros(rt.WS_STATS_FLOATS // 8))
    set_(rt, "stream", lambda: 0)
    set_(rt, "synchronize", lambda: None)
    return fake

"""TEST INFRASTRUCTURE: a stand-in for the four critic entry points of libb200voc.so that evaluates the SAME
index arithmetic as csrc/disc.cu (flattened (row, column) positions, in_batch_stride / in_valid reads, output
offsets) with numpy on host pointers.  It lets the CPU suite check the host-side layer walker of
b200voc/discriminators.py -- pointer offsets of the time chunks, the padded period view, the pooled scales --
against the oracle without a GPU.  It is never importable from the product package."""
import ctypes as C

import numpy as np


def _arr(ptr: int, n: int) -> np.ndarray:
    return np.ctypeslib.as_array((C.c_float * int(n)).from_address(int(ptr)))


class FakeCriticLib:
    def __init__(self):
        self.conv_calls = []

    def b200voc_device_supported(self, dev):
        return 0

    def b200voc_last_error_string(self):
        return b"fake"

    def b200voc_disc_conv_tc_supported(self, Cin, Cout, K, stride, P):
        return 0          # the stand-in has no tensor-core path: the walker must fall back to b200voc_disc_conv

    def b200voc_disc_conv_out_len(self, Lin, K, stride, pad):
        if Lin <= 0 or K <= 0 or stride <= 0 or pad < 0 or Lin + 2 * pad < K:
            return 0
        return (Lin + 2 * pad - K) // stride + 1

    def b200voc_disc_conv(self, x, w, bias, B, Cin, Cout, Lin, P, K, stride, pad, in_batch_stride, in_valid, slope,
                          y_pre, y_act, stream):
        Lout = self.b200voc_disc_conv_out_len(Lin, K, stride, pad)
        bs = in_batch_stride if in_batch_stride > 0 else Cin * Lin * P
        valid = in_valid if in_valid > 0 else Lin * P
        chan = Lin * P
        self.conv_calls.append((B, Cin, Cout, Lin, P, K, stride, pad, bs, valid))
        wv = _arr(w, Cout * Cin * K).reshape(Cout, Cin * K)
        bv = _arr(bias, Cout)
        pos = np.arange(Lout * P)
        lo = pos // P
        col = pos - lo * P
        li = (lo * stride - pad)[None, :] + np.arange(K)[:, None]                 # [K, pos]
        idx = li * P + col[None, :]
        ok = (li >= 0) & (li < Lin) & (idx < valid)
        out = np.empty((B, Cout, Lout * P), dtype=np.float32)
        for b in range(B):
            cols = np.zeros((Cin, K, Lout * P), dtype=np.float32)
            for ci in range(Cin):
                base = b * bs + ci * chan
                span = _arr(x + 4 * base, max(int(idx[ok].max()) + 1, 1)) if ok.any() else None
                if span is not None:
                    cols[ci][ok] = span[idx[ok]]
            out[b] = wv @ cols.reshape(Cin * K, Lout * P) + bv[:, None]
        if y_pre:
            _arr(y_pre, out.size)[:] = out.reshape(-1)
        if y_act:
            _arr(y_act, out.size)[:] = np.where(out > 0, out, np.float32(slope) * out).reshape(-1)
        return 0

    def b200voc_spectral_norm_weight(self, w_orig, u, v, rows, cols, w_out, sigma_out, stream):
        W = _arr(w_orig, rows * cols).reshape(rows, cols)
        sigma = np.float32(_arr(u, rows) @ (W @ _arr(v, cols)))
        _arr(sigma_out, 1)[0] = sigma
        _arr(w_out, rows * cols)[:] = (W / sigma).reshape(-1)
        return 0

    def b200voc_spectral_norm_train(self, w_orig, u, v, rows, cols, eps, w_out, sigma_out, scratch, stream):
        """one power iteration, u / v updated IN PLACE (csrc/disc.cu spectral_norm_train_launch)"""
        W = _arr(w_orig, rows * cols).reshape(rows, cols).astype(np.float64)
        uu, vv = _arr(u, rows), _arr(v, cols)
        t = (W.T @ uu.astype(np.float64)).astype(np.float32)
        vv[:] = t / max(float(np.sqrt((t.astype(np.float64) ** 2).sum())), eps)
        s = (W @ vv.astype(np.float64)).astype(np.float32)
        tot = float((s.astype(np.float64) ** 2).sum())
        den = max(float(np.float32(np.sqrt(tot))), eps)
        uu[:] = s / np.float32(den)
        sigma = np.float32(tot / den)
        _arr(sigma_out, 1)[0] = sigma
        _arr(w_out, rows * cols)[:] = (_arr(w_orig, rows * cols) / sigma)
        return 0

    def b200voc_avg_pool1d_k4s2p1(self, x, rows, Lin, y, stream):
        Lout = (Lin + 2 - 4) // 2 + 1
        xv = np.pad(_arr(x, rows * Lin).reshape(rows, Lin), ((0, 0), (1, 2 * Lout + 2 - Lin)))
        out = sum(xv[:, k:k + 2 * Lout:2] for k in range(4)) * np.float32(0.25)
        _arr(y, rows * Lout)[:] = out.astype(np.float32).reshape(-1)
        return 0

    # ------------------------------------------------------------------ backward (csrc/disc_bwd.cu conventions)
    def b200voc_disc_conv_dgrad_tc_supported(self, cin, cout, k, stride, P, pad):
        return 0

    def b200voc_disc_conv_wgrad_tc_supported(self, B, cin, cout, Lin, k, stride, P, pad):
        return 0

    def b200voc_disc_conv_wgrad_scratch_bytes(self, B, cin, cout, Lin, P, K, stride, pad):
        return 0

    def b200voc_spectral_norm_bwd_scratch_bytes(self):
        return 2048

    def b200voc_disc_lrelu_bwd(self, y, gy, ga, gn, slope, n, g_out, stream):
        a = np.zeros(n, np.float32)
        if ga:
            a += _arr(ga, n)
        if gn:
            a += _arr(gn, n)
        r = a * np.where(_arr(y, n) > 0, np.float32(1), np.float32(slope)) if (ga or gn) else np.zeros(n, np.float32)
        if gy:
            r = r + _arr(gy, n)
        _arr(g_out, n)[:] = r
        return 0

    def b200voc_disc_bias_grad(self, g, B, cout, per_c, db, stream):
        _arr(db, cout)[:] = _arr(g, B * cout * per_c).reshape(B, cout, per_c).sum(axis=(0, 2))
        return 0

    def _gather_index(self, Lin, P, K, stride, pad, Lout):
        li = np.arange(Lout)[:, None] * stride - pad + np.arange(K)[None, :]                 # [Lout, K]
        idx = li[:, :, None] * P + np.arange(P)[None, None, :]                                # [Lout, K, P]
        return li, idx

    def b200voc_disc_conv_dgrad(self, g, w, B, cin, cout, Lin, P, K, stride, pad, in_batch_stride, in_valid, accumulate,
                                dx, stream):
        Lout = self.b200voc_disc_conv_out_len(Lin, K, stride, pad)
        sb = in_batch_stride or cin * Lin * P
        valid = in_valid or Lin * P
        gv = _arr(g, B * cout * Lout * P).reshape(B, cout, Lout, P)
        wv = _arr(w, cout * cin * K).reshape(cout, cin, K)
        li, idx = self._gather_index(Lin, P, K, stride, pad, Lout)
        ok = (li[:, :, None] >= 0) & (li[:, :, None] < Lin) & (idx < valid)
        contrib = np.einsum("bolp,oik->bilkp", gv, wv)                                        # [B, cin, Lout, K, P]
        for b in range(B):
            for ci in range(cin):
                row_len = min(valid, Lin * P)
                row = _arr(dx + 4 * (b * sb + ci * Lin * P), row_len)
                acc = np.zeros(Lin * P, np.float32)
                np.add.at(acc, np.where(ok, idx, 0).ravel(), np.where(ok, contrib[b, ci], 0).ravel().astype(np.float32))
                row[:] = (row if accumulate else 0) + acc[:row_len]
        return 0

    def b200voc_disc_conv_wgrad(self, x, g, B, cin, cout, Lin, P, K, stride, pad, in_batch_stride, in_valid, dw, scratch,
                                stream):
        Lout = self.b200voc_disc_conv_out_len(Lin, K, stride, pad)
        sb = in_batch_stride or cin * Lin * P
        valid = in_valid or Lin * P
        gv = _arr(g, B * cout * Lout * P).reshape(B, cout, Lout, P)
        li, idx = self._gather_index(Lin, P, K, stride, pad, Lout)
        ok = (li[:, :, None] >= 0) & (li[:, :, None] < Lin) & (idx < valid)
        cols = np.zeros((B, cin, Lout, K, P), np.float32)
        for b in range(B):
            for ci in range(cin):
                row = _arr(x + 4 * (b * sb + ci * Lin * P), min(valid, Lin * P))
                cols[b, ci] = np.where(ok, row[np.where(ok, idx, 0)], 0)
        _arr(dw, cout * cin * K).reshape(cout, cin, K)[:] = np.einsum("bolp,bilkp->oik", gv, cols)
        return 0

    def b200voc_avg_pool1d_k4s2p1_bwd(self, gy, rows, Lin, dx, stream):
        Lout = (Lin - 2) // 2 + 1
        g = _arr(gy, rows * Lout).reshape(rows, Lout)
        pad = np.zeros((rows, Lin + 4), np.float32)
        for k in range(4):
            pad[:, k:k + 2 * Lout:2] += 0.25 * g
        _arr(dx, rows * Lin).reshape(rows, Lin)[:] = pad[:, 1:1 + Lin]
        return 0

    def b200voc_spectral_norm_bwd(self, dw, w, u, v, sigma, rows, cols, out, scratch, stream):
        dwv, wv = _arr(dw, rows * cols).reshape(rows, cols), _arr(w, rows * cols).reshape(rows, cols)
        dot = float((dwv.astype(np.float64) * wv).sum())
        _arr(out, rows * cols).reshape(rows, cols)[:] = (dwv - np.float32(dot) * np.outer(_arr(u, rows), _arr(v, cols))) \
            / _arr(sigma, 1)[0]
        return 0

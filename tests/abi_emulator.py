"""TEST INFRASTRUCTURE: a numpy restatement of the C ABI in include/demucs_b200.h.

It lets the CPU test-suite exercise the product's HOST logic -- weight packing, implicit-GEMM
descriptors, buffer plumbing, the apply_model batcher and the multi-rank sharder -- without a
GPU, by standing in for libdemucs_b200.so behind ``demucs_b200._lib.TEST_HOOK``.  It is written
against the header's documented semantics (not against the CUDA sources) and is never imported by
the product.  The CUDA kernels themselves are checked by the ``-m gpu`` tests.
"""
import contextlib
import ctypes as C
import math

import numpy as np

from demucs_b200 import _lib


def f32(p, n):
    return np.ctypeslib.as_array(C.cast(C.c_void_p(p), C.POINTER(C.c_float)), shape=(int(n),))


def f64(p, n):
    return np.ctypeslib.as_array(C.cast(C.c_void_p(p), C.POINTER(C.c_double)), shape=(int(n),))


def gelu(x):
    from scipy.special import erf
    return (0.5 * x * (1.0 + erf(x / np.sqrt(2.0)))).astype(np.float32)


def sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(np.float32)


def bd_stft_cac(mix, window, twiddle, spec, stats, B, A, L, stream):
    T = (L + 1023) // 1024
    x = f32(mix, B * A * L).reshape(B, A, L)
    win = f32(window, 4096)
    idx = np.arange(T)[:, None] * 1024 + np.arange(4096)[None, :] - 1536
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= L, 2 * (L - 1) - idx, idx)
    z = np.fft.rfft(x[:, :, idx] * win, axis=-1)[..., :2048] / 64.0      # [B,A,T,F]
    out = f32(spec, B * T * 2048 * 4).reshape(B, T, 2048, 2, 2)
    out[..., 0] = z.real.transpose(0, 2, 3, 1)
    out[..., 1] = z.imag.transpose(0, 2, 3, 1)
    st = f64(stats, 4 * B).reshape(B, 4)
    o64 = out.reshape(B, -1).astype(np.float64)
    x64 = x.reshape(B, -1).astype(np.float64)
    st[:, 0] += o64.sum(1); st[:, 1] += (o64 ** 2).sum(1)
    st[:, 2] += x64.sum(1); st[:, 3] += (x64 ** 2).sum(1)


def bd_finalize_item_norm(stats, norm, B, n_freq, n_time, stream):
    st = f64(stats, 4 * B).reshape(B, 2, 2)
    nm = f32(norm, 8 * B).reshape(B, 2, 4)
    for which, n in ((0, n_freq), (1, n_time)):
        mean = st[:, which, 0] / n
        var = (st[:, which, 1] - st[:, which, 0] * mean) / (n - 1.0)
        sd = np.sqrt(np.maximum(var, 0)).astype(np.float32)
        nm[:, which, 0] = mean; nm[:, which, 1] = sd
        nm[:, which, 2] = np.float32(1.0) / (np.float32(1e-5) + sd); nm[:, which, 3] = 0


def bd_istft_frames(spec, norm, window, twiddle, frames, B, S, T, stream):
    x = f32(spec, B * T * 2048 * 4 * S).reshape(B, T, 2048, S, 2, 2)
    nm = f32(norm, 8 * B).reshape(B, 8)
    x = x * nm[:, 1].reshape(B, 1, 1, 1, 1, 1) + nm[:, 0].reshape(B, 1, 1, 1, 1, 1)
    z = (x[..., 0] + 1j * x[..., 1]).transpose(0, 3, 4, 1, 2)             # [B,S,2,T,F]
    z = np.concatenate([z, np.zeros_like(z[..., :1])], axis=-1)
    fr = np.fft.irfft(z, n=4096, axis=-1) * 64.0 * f32(window, 4096) / 1.5
    f32(frames, B * S * 2 * T * 4096)[:] = fr.astype(np.float32).reshape(-1)


def bd_istft_ola(spec, norm, window, twiddle, xt, out, B, S, T, Lseg, Lout, stream):
    x = f32(spec, B * T * S * 2048 * 4).reshape(B, T, S, 2048, 2, 2)
    nm = f32(norm, 8 * B).reshape(B, 8)
    x = x * nm[:, 1].reshape(B, 1, 1, 1, 1, 1) + nm[:, 0].reshape(B, 1, 1, 1, 1, 1)
    z = (x[..., 0] + 1j * x[..., 1]).transpose(0, 2, 4, 1, 3)             # [B,S,2,T,F]
    z = np.concatenate([z, np.zeros_like(z[..., :1])], axis=-1)
    fr = (np.fft.irfft(z, n=4096, axis=-1) * 64.0 * f32(window, 4096) / 1.5).astype(np.float32)
    acc = np.zeros((B, S, 2, 1024 * (T - 1) + 4096), np.float32)
    for t in range(T):
        acc[..., t * 1024: t * 1024 + 4096] += fr[:, :, :, t]
    res = acc[..., 1536: 1536 + Lout].copy()
    if xt:
        xv = f32(xt, B * Lseg * 2 * S).reshape(B, Lseg, S, 2)[:, :Lout].transpose(0, 2, 3, 1)
        res += xv * nm[:, 5].reshape(B, 1, 1, 1) + nm[:, 4].reshape(B, 1, 1, 1)
    f32(out, B * 2 * S * Lout)[:] = res.reshape(-1)


def bd_ola_combine(frames, xt, norm, out, B, S, T, Lseg, Lout, stream):
    fr = f32(frames, B * S * 2 * T * 4096).reshape(B, 2 * S, T, 4096)
    acc = np.zeros((B, 2 * S, 1024 * (T - 1) + 4096), np.float32)
    for t in range(T):
        acc[..., t * 1024: t * 1024 + 4096] += fr[:, :, t]
    res = acc[..., 1536: 1536 + Lout].copy()
    if xt:
        nm = f32(norm, 8 * B).reshape(B, 8)
        x = f32(xt, B * Lseg * 2 * S).reshape(B, Lseg, 2 * S)[:, :Lout].transpose(0, 2, 1)
        res += x * nm[:, 5].reshape(B, 1, 1) + nm[:, 4].reshape(B, 1, 1)
    f32(out, B * 2 * S * Lout)[:] = res.reshape(-1)


def bd_conv_gemm(dref, stream):
    d = dref._obj
    M, N, K, Cin = d.M, d.N, d.K, d.Cin
    m = np.arange(M, dtype=np.int64)
    i0 = m % d.I0
    t = m // d.I0
    i1 = t % d.I1
    b = t // d.I1
    w = f32(d.w, N * K).reshape(N, K)
    A = np.zeros((M, K), np.float32)
    ci = np.arange(Cin, dtype=np.int64) * d.xs_c
    for tap in range(d.taps):
        j1 = i1 * d.m1 + d.d1[tap]
        j0 = i0 * d.m0 + d.d0[tap]
        ok = (j1 >= 0) & (j1 < d.J1) & (j0 >= 0) & (j0 < d.J0)
        base = b * d.xs_b + j1 * d.xs_1 + j0 * d.xs_0
        idx = base[ok][:, None] + ci[None, :]
        if idx.size:
            x = f32(d.x, int(idx.max()) + 1)
            vals = x[idx]
            if d.a_mode == _lib.A_GN_GELU:
                sl = (m // d.stat_div) * d.stat_mul + (m % d.stat_mod)
                st = f32(d.a_stats, 2 * (int(sl.max()) + 1)).reshape(-1, 2)[sl[ok]]
                g, be = f32(d.a_gamma, Cin), f32(d.a_beta, Cin)
                vals = gelu((vals - st[:, :1]) * st[:, 1:2] * g + be)
            elif d.a_mode == _lib.A_ITEM_AFFINE:
                nb = int(b.max()) + 1
                st = f32(d.a_stats, nb * d.a_stats_stride).reshape(nb, d.a_stats_stride)[b[ok]]
                vals = (vals - st[:, :1]) * st[:, 2:3]
            A[ok, tap * Cin:(tap + 1) * Cin] = vals
    v = A @ w.T
    if d.bias:
        v = v + f32(d.bias, N)
    if d.e_stats:
        sl = (m // d.stat_div) * d.stat_mul + (m % d.stat_mod)
        st = f32(d.e_stats, 2 * (int(sl.max()) + 1)).reshape(-1, 2)[sl]
        v = (v - st[:, :1]) * st[:, 1:2] * f32(d.e_gamma, N) + f32(d.e_beta, N)
    if d.act == _lib.ACT_GELU:
        v = gelu(v)
    elif d.act == _lib.ACT_GLU:
        v = v[:, 0::2] * sigmoid(v[:, 1::2])
    nout = v.shape[1]
    n = np.arange(nout, dtype=np.int64)
    if d.convt:
        cout = N // 4
        r, co = n // cout, n % cout
        o0 = 4 * i0[:, None] + r[None, :] - (2 if d.convt == 1 else 0)
        ok = (o0 >= 0) & (o0 < d.O0)
        colofs = co if d.oc_split == 0 else (co // d.oc_split) * d.oc_stride + co % d.oc_split
        oidx = (b * d.os_b + i1 * d.os_1)[:, None] + o0 * d.os_0 + colofs[None, :]
    else:
        ok = np.ones((M, nout), bool)
        colofs = n if d.oc_split == 0 else (n // d.oc_split) * d.oc_stride + n % d.oc_split
        oidx = (b * d.os_b + i1 * d.os_1 + i0 * d.os_0)[:, None] + colofs[None, :]
    if d.rowbias:
        rb = f32(d.rowbias, d.rowbias_period * nout).reshape(d.rowbias_period, nout)
        v = v + rb[m % d.rowbias_period]
    nmax = int(oidx[ok].max()) + 1
    if d.resid:
        sc = f32(d.scale, nout) if d.scale else np.ones(nout, np.float32)
        v = np.where(ok, f32(d.resid, nmax)[np.where(ok, oidx, 0)] + sc * v, 0)
    if d.addend:
        v = v + np.where(ok, f32(d.addend, nmax)[np.where(ok, oidx, 0)], 0)
    v = v.astype(np.float32)
    if d.out:
        f32(d.out, nmax)[oidx[ok]] = v[ok]
    if d.stats_out:
        sl = (m // d.stat_div) * d.stat_mul + (m % d.stat_mod)
        slabs = int(sl.max()) + 1
        st = f64(d.stats_out, 2 * slabs).reshape(slabs, 2)
        v64 = np.where(ok, v, 0).astype(np.float64)
        np.add.at(st[:, 0], sl, v64.sum(1))
        np.add.at(st[:, 1], sl, (v64 ** 2).sum(1))


def bd_finalize_group_stats(sums, mean_rstd, slabs, count, stream):
    s = f64(sums, 2 * slabs).reshape(slabs, 2)
    o = f32(mean_rstd, 2 * slabs).reshape(slabs, 2)
    mean = s[:, 0] / count
    var = np.maximum(s[:, 1] / count - mean * mean, 0)
    o[:, 0] = mean
    o[:, 1] = 1.0 / np.sqrt(var + 1e-5)
    s[:] = 0.0          # the accumulators are handed back cleared


def bd_dconv_tail(x, u, mr, gamma, beta, scale, M, Cc, rows_per_item, spi, stream):
    xv = f32(x, M * Cc).reshape(M, Cc)
    uv = f32(u, M * 2 * Cc).reshape(M, 2 * Cc)
    m = np.arange(M)
    slab = (m // rows_per_item) * spi + (m % spi)
    st = f32(mr, 2 * (int(slab.max()) + 1)).reshape(-1, 2)[slab]
    g = (uv - st[:, :1]) * st[:, 1:2] * f32(gamma, 2 * Cc) + f32(beta, 2 * Cc)
    xv += f32(scale, Cc) * (g[:, 0::2] * sigmoid(g[:, 1::2]))


def _dconv_u(h, ldh, hid, mr1, g1, be1, w2t, b2, M, Cc, rpi, spi):
    hv = f32(h, M * ldh).reshape(M, ldh)[:, :hid]
    m = np.arange(M)
    slab = (m // rpi) * spi + (m % spi)
    st = f32(mr1, 2 * (int(slab.max()) + 1)).reshape(-1, 2)[slab]
    g = gelu((hv - st[:, :1]) * st[:, 1:2] * f32(g1, hid) + f32(be1, hid))
    u = g @ f32(w2t, hid * 2 * Cc).reshape(hid, 2 * Cc) + f32(b2, 2 * Cc)
    return u.astype(np.float32), slab


def bd_encoder_conv0(x, channel_major, norm, norm_stride, w, bias, out, B, I1, Io, Jin, cin, cout, math, stream):
    nv = f32(norm, (B - 1) * norm_stride + 3)
    wv = f32(w, cout * 8 * cin).reshape(cout, 8, cin)
    ov = f32(out, B * I1 * Io * cout).reshape(B, I1, Io, cout)
    if channel_major:
        xv = f32(x, B * cin * Jin).reshape(B, 1, cin, Jin).transpose(0, 1, 3, 2)      # [B, 1, Jin, cin]
    else:
        xv = f32(x, B * I1 * Jin * cin).reshape(B, I1, Jin, cin)
    for b in range(B):
        mean, rstd = nv[b * norm_stride], nv[b * norm_stride + 2]
        xn = np.zeros((I1, 4 * Io + 8, cin), np.float32)                           # window of i0: rows 4*i0 .. 4*i0+7
        hi = min(Jin, 4 * Io + 6)
        xn[:, 2:2 + hi] = (xv[b, :, :hi] - mean) * rstd
        acc = np.tile(f32(bias, cout), (I1, Io, 1)).astype(np.float32)
        for tap in range(8):
            acc += xn[:, tap:tap + 4 * Io:4] @ wv[:, tap].T
        ov[b] = gelu(acc)


def bd_dconv_conv3(x, w1, b1, h, ldh, sums1, M, Cc, hid, rpi, spi, dil, math, stream):
    xv = f32(x, M * Cc).reshape(M, Cc)
    w = f32(w1, hid * 3 * Cc).reshape(hid, 3, Cc)
    m = np.arange(M)
    t = (m % rpi) // spi
    T = rpi // spi
    acc = np.tile(f32(b1, hid), (M, 1)).astype(np.float32)
    for tap in range(3):
        tt = t + (tap - 1) * dil
        ok = (tt >= 0) & (tt < T)
        src = np.where(ok, m + (tap - 1) * dil * spi, 0)
        acc += np.where(ok[:, None], xv[src], 0.0) @ w[:, tap].T
    hv = f32(h, M * ldh).reshape(M, ldh)
    hv[:, :hid] = acc
    hv[:, hid:] = 0.0
    slab = (m // rpi) * spi + (m % spi)
    st = f64(sums1, 2 * (int(slab.max()) + 1)).reshape(-1, 2)
    a64 = acc.astype(np.float64)
    np.add.at(st[:, 0], slab, a64.sum(1))
    np.add.at(st[:, 1], slab, (a64 ** 2).sum(1))


def bd_dconv_expand_stats(h, ldh, hid, mr1, g1, be1, w2t, b2, sums2, gram_ws, M, Cc, rpi, spi, stream):
    u, slab = _dconv_u(h, ldh, hid, mr1, g1, be1, w2t, b2, M, Cc, rpi, spi)
    st = f64(sums2, 2 * (int(slab.max()) + 1)).reshape(-1, 2)
    u64 = u.astype(np.float64)
    np.add.at(st[:, 0], slab, u64.sum(1))
    np.add.at(st[:, 1], slab, (u64 ** 2).sum(1))


def bd_dconv_expand_update(h, ldh, hid, mr1, g1, be1, w2t, b2, mr2, g2, be2, scale, x, M, Cc, rpi, spi, math_, stream):
    u, slab = _dconv_u(h, ldh, hid, mr1, g1, be1, w2t, b2, M, Cc, rpi, spi)
    st = f32(mr2, 2 * (int(slab.max()) + 1)).reshape(-1, 2)[slab]
    v = (u - st[:, :1]) * st[:, 1:2] * f32(g2, 2 * Cc) + f32(be2, 2 * Cc)
    xv = f32(x, M * Cc).reshape(M, Cc)
    xv += f32(scale, Cc) * (v[:, 0::2] * sigmoid(v[:, 1::2]))


def bd_gn_gelu_apply(h, mr, gamma, beta, M, Cc, rows_per_item, spi, stream):
    hv = f32(h, M * Cc).reshape(M, Cc)
    m = np.arange(M)
    slab = (m // rows_per_item) * spi + (m % spi)
    st = f32(mr, 2 * (int(slab.max()) + 1)).reshape(-1, 2)[slab]
    hv[:] = gelu((hv - st[:, :1]) * st[:, 1:2] * f32(gamma, Cc) + f32(beta, Cc))


def bd_layer_norm(x, y, gamma, beta, pos, period, M, Cc, y_bf16, stream):
    assert not y_bf16, "the emulator keeps fp32 tensors"
    xv = f32(x, M * Cc).reshape(M, Cc).astype(np.float64)
    mean = xv.mean(1, keepdims=True)
    var = xv.var(1, keepdims=True)
    o = (xv - mean) / np.sqrt(var + 1e-5) * f32(gamma, Cc) + f32(beta, Cc)
    if pos:
        o = o + f32(pos, period * Cc).reshape(period, Cc)[np.arange(M) % period]
    f32(y, M * Cc)[:] = o.astype(np.float32).reshape(-1)


def bd_item_stats(x, sums, B, n, stream):
    v = f32(x, B * n).reshape(B, n).astype(np.float64)
    st = f64(sums, 2 * B).reshape(B, 2)
    st[:, 0] += v.sum(1)
    st[:, 1] += (v ** 2).sum(1)


def bd_group_norm_apply(x, mr, gamma, beta, B, rows, Cc, stream):
    xv = f32(x, B * rows * Cc).reshape(B, rows, Cc)
    st = f32(mr, 2 * B).reshape(B, 2)
    xv[:] = (xv - st[:, 0].reshape(B, 1, 1)) * st[:, 1].reshape(B, 1, 1) * f32(gamma, Cc) + f32(beta, Cc)


def _strided(p, B, T, ld, ncol):
    a = f32(p, (B * T - 1) * ld + ncol)
    return np.lib.stride_tricks.as_strided(a, shape=(B, T, ncol), strides=(T * ld * 4, ld * 4, 4))


def bd_attention_workspace(B, H, Tq, Tk, math_):
    D, Tkp = 64 * H, (Tk + 3) // 4 * 4
    if math_ == _lib.MATH_TF32:
        return B * D * Tkp
    if math_ == _lib.MATH_TF32X3:
        return 2 * B * D * (Tkp + Tq + Tk)
    return 0


def bd_attention(q, k, v, o, B, H, Tq, Tk, ldq, ldk, ldv, ldo, math_, ws, stream):
    D = 64 * H
    qv, kv, vv = _strided(q, B, Tq, ldq, D), _strided(k, B, Tk, ldk, D), _strided(v, B, Tk, ldv, D)
    ov = _strided(o, B, Tq, ldo, D)
    for h in range(H):
        sl = slice(64 * h, 64 * h + 64)
        s = np.einsum("bqd,bkd->bqk", qv[..., sl], kv[..., sl]).astype(np.float64) / 8.0
        s = np.exp(s - s.max(-1, keepdims=True))
        p = s / s.sum(-1, keepdims=True)
        ov[..., sl] = np.einsum("bqk,bkd->bqd", p, vv[..., sl]).astype(np.float32)


def bd_overlap_add(segs, weight, out, seg_first, nseg_local, nseg, rows, valid, seg_len, stride, length, out_ld,
                   out_shift, n_begin, n_end, row_alpha, alpha, accumulate, stream):
    sv = f32(segs, nseg_local * rows * valid).reshape(nseg_local, rows, valid)
    w = f32(weight, seg_len)
    n_begin, n_end = max(n_begin, out_shift), min(n_end, length)
    if n_end <= n_begin:
        return
    num = np.zeros((rows, length), np.float32)
    den = np.zeros(length, np.float32)
    for i in range(seg_first, seg_first + nseg_local):
        off = i * stride
        n_i = min(length - off, seg_len)
        lead = (valid - n_i) // 2
        num[:, off:off + n_i] += w[:n_i] * sv[i - seg_first][:, lead:lead + n_i]
        den[off:off + n_i] += w[:n_i]
    ra = f32(row_alpha, rows)[:, None] if row_alpha else 1.0
    res = (num[:, n_begin:n_end] / den[n_begin:n_end]) * np.float32(alpha) * ra
    ov = f32(out, (rows - 1) * out_ld + n_end - out_shift)
    ov = np.lib.stride_tricks.as_strided(ov, shape=(rows, n_end - out_shift), strides=(out_ld * 4, 4))
    if accumulate:
        ov[:, n_begin - out_shift:] += res
    else:
        ov[:, n_begin - out_shift:] = res


def bd_gather_segments(track, batch, B, Cc, track_len, offset0, length, seg_first, nseg_batch, seg_len, stride, valid, stream):
    tr = f32(track, B * Cc * track_len).reshape(B * Cc, track_len)
    out = f32(batch, nseg_batch * B * Cc * valid).reshape(nseg_batch, B * Cc, valid)
    for j in range(nseg_batch):
        i = seg_first + j
        n_i = min(length - i * stride, seg_len)
        start = offset0 + i * stride - (valid - n_i) // 2
        lo, hi = max(0, start), min(track_len, start + valid)
        out[j] = 0
        if hi > lo:
            out[j][:, lo - start:hi - start] = tr[:, lo:hi]


def bd_gn_stats(x, sums, B, rows, Cc, G, stream):
    v = f32(x, B * rows * Cc).reshape(B, rows, G, Cc // G).astype(np.float64)
    st = f64(sums, 2 * B * G).reshape(B, G, 2)
    st[..., 0] += v.sum(axis=(1, 3))
    st[..., 1] += (v ** 2).sum(axis=(1, 3))


def bd_gn_act(x, y, mr, gamma, beta, addend, B, rows_in, row0, rows_out, Cc, G, act, y_item_stride, stream):
    v = f32(x, B * rows_in * Cc).reshape(B, rows_in, Cc)[:, row0:row0 + rows_out]
    m = f32(mr, 2 * B * G).reshape(B, G, 2)
    mean = np.repeat(m[:, :, 0], Cc // G, axis=1)[:, None, :]
    rstd = np.repeat(m[:, :, 1], Cc // G, axis=1)[:, None, :]
    v = (v - mean) * rstd * f32(gamma, Cc) + f32(beta, Cc)
    if act == _lib.ACT_GELU:
        v = gelu(v)
    elif act == _lib.ACT_GLU:
        v = v[..., :Cc // 2] * sigmoid(v[..., Cc // 2:])
    Co = v.shape[-1]
    out = f32(y, (B - 1) * y_item_stride + rows_out * Co)
    add = f32(addend, (B - 1) * y_item_stride + rows_out * Co) if addend else None
    for b in range(B):
        sl = slice(b * y_item_stride, b * y_item_stride + rows_out * Co)
        res = v[b].astype(np.float32).reshape(-1)
        out[sl] = res + add[sl] if add is not None else res


def bd_lstm_frame(x, frames, B, T, Cc, nf, width, stride, stream):
    v = f32(x, B * T * Cc).reshape(B, T, Cc)
    pad = np.zeros((B, (nf - 1) * stride + width, Cc), np.float32)
    pad[:, :T] = v
    out = f32(frames, B * nf * width * Cc).reshape(B, nf, width, Cc)
    for k in range(nf):
        out[:, k] = pad[:, k * stride: k * stride + width]


def bd_lstm_unframe_add(frames, skip, out, B, T, Cc, nf, width, stride, stream):
    fr = f32(frames, B * nf * width * Cc).reshape(B, nf, width, Cc)
    t = np.arange(T)
    limit = stride // 2
    k = np.clip(np.where(t < limit, 0, (t - limit) // stride), 0, nf - 1)
    f32(out, B * T * Cc).reshape(B, T, Cc)[:] = fr[:, k, t - k * stride] + f32(skip, B * T * Cc).reshape(B, T, Cc)


def bd_lstm_bidir(pre, whhT, out, ws, N, T, H, stream):
    p = f32(pre, N * T * 8 * H).reshape(N, T, 2, 4 * H)
    w = f32(whhT, 2 * H * 4 * H).reshape(2, H, 4 * H)
    o = f32(out, N * T * 2 * H).reshape(N, T, 2, H)
    for d in range(2):
        h = np.zeros((N, H), np.float32)
        c = np.zeros((N, H), np.float32)
        for step in range(T):
            t = T - 1 - step if d else step
            g = p[:, t, d] + h @ w[d]
            i, f, gg, oo = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
            c = sigmoid(f) * c + sigmoid(i) * np.tanh(gg)
            h = (sigmoid(oo) * np.tanh(c)).astype(np.float32)
            o[:, t, d] = h


def bd_local_state(qkc, dq, out, N, T, D, heads, stream):
    v = f32(qkc, N * T * 3 * D).reshape(N, T, 3, heads, D // heads)
    q, k, c = v[:, :, 0], v[:, :, 1], v[:, :, 2]
    dots = np.einsum("nthc,nshc->nhts", k, q) / np.sqrt(D // heads)
    dec = sigmoid(f32(dq, N * T * heads * 4).reshape(N, T, heads, 4)) / 2                    # [n, s, h, f]
    slope = (dec * (np.arange(1, 5, dtype=np.float32) / 2.0)).sum(-1)                        # [n, s, h]
    idx = np.arange(T)
    dots = dots - np.abs(idx[:, None] - idx[None, :])[None, None] * slope.transpose(0, 2, 1)[:, :, None, :]
    dots[:, :, idx, idx] = -100.0
    w = np.exp(dots - dots.max(axis=2, keepdims=True))
    w = w / w.sum(axis=2, keepdims=True)
    f32(out, N * T * D).reshape(N, T, heads, D // heads)[:] = np.einsum("nhts,nthc->nshc", w, c).astype(np.float32)


def _mix_channels(x, src, dst):
    if src == dst or (src > dst and dst != 1):
        return x[:, :dst]
    if src == 1:
        return np.repeat(x, dst, axis=1)
    return x.mean(axis=1, keepdims=True).astype(np.float32)


def bd_convert_channels(x, y, items, src, dst, length, stream):
    f32(y, items * dst * length).reshape(items, dst, length)[:] = _mix_channels(
        f32(x, items * src * length).reshape(items, src, length), src, dst)


def bd_resample_frac(x, y, kernel, items, src, dst, Lin, Lout, old_sr, new_sr, width, stream):
    xin = _mix_channels(f32(x, items * src * Lin).reshape(items, src, Lin), src, dst)
    klen = 2 * width + old_sr
    k = f32(kernel, new_sr * klen).reshape(new_sr, klen)
    out = f32(y, items * dst * Lout).reshape(items, dst, Lout)
    o = np.arange(Lout)
    j, i = o // new_sr, o % new_sr
    idx = np.clip(j[:, None] * old_sr - width + np.arange(klen)[None, :], 0, Lin - 1)
    out[:] = np.einsum("bcok,ok->bco", xin[:, :, idx], k[i]).astype(np.float32)


def bd_absmax(x, peak, n, stream):
    f32(peak, 1)[0] = np.abs(f32(x, n)).max()


def bd_clip_pcm(x, out, channels, frames, mode, peak, bits, stream):
    v = f32(x, channels * frames).reshape(channels, frames).copy()
    if mode == 1:
        v = v * (np.float32(1.0) / max(np.float32(1.01) * f32(peak, 1)[0], np.float32(1.0)))
    elif mode == 2:
        v = np.clip(v, -0.99, 0.99)
    elif mode == 3:
        v = np.tanh(v)
    v = v.T
    if bits == 32:
        f32(out, channels * frames).reshape(frames, channels)[:] = v
        return
    v = np.clip(v, -1, 1)
    if bits == 16:
        dst = np.ctypeslib.as_array(C.cast(C.c_void_p(out), C.POINTER(C.c_int16)), shape=(frames * channels,))
        dst.reshape(frames, channels)[:] = np.trunc(v * np.float32(32767.0)).astype(np.int16)
    else:
        dst = np.ctypeslib.as_array(C.cast(C.c_void_p(out), C.POINTER(C.c_int32)), shape=(frames * channels,))
        dst.reshape(frames, channels)[:] = np.trunc(v * np.float32(8388607.0)).astype(np.int32)


TABLE = {k: v for k, v in globals().items() if k.startswith("bd_")}
CALLS = []


def hook(name, *args):
    CALLS.append(name)
    args = [a.value if isinstance(a, C.c_void_p) else a for a in args]
    return TABLE[name](*args)


@contextlib.contextmanager
def emulated_abi():
    """Route demucs_b200's kernel calls to the numpy restatement for the duration of the block."""
    prev = _lib.TEST_HOOK
    _lib.TEST_HOOK = hook
    try:
        yield
    finally:
        _lib.TEST_HOOK = prev

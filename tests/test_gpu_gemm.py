"""GPU: the implicit-GEMM kernel family (fp32 CUDA-core arm and tcgen05 TF32 arm) against torch
on the same operands, through the C ABI, for each fused-epilogue flavour the engine uses."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from _fixtures import rel_l2
from demucs_b200 import _lib
from demucs_b200._lib import GemmDesc, ptr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run_plain(M, N, K, math_mode, bias=True, act=_lib.ACT_NONE, resid=False, stats_rows=0, lda=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    lda = lda or K
    xfull = torch.randn(M, lda, generator=g).to(DEV)
    x = xfull[:, :K]
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV) if bias else None
    n_out = N // 2 if act == _lib.ACT_GLU else N
    r = torch.randn(M, n_out, generator=g).to(DEV) if resid else None
    sc = torch.randn(n_out, generator=g).to(DEV) if resid else None
    out = torch.full((M, n_out), float("nan"), device=DEV)
    slabs = M // stats_rows if stats_rows else 0
    sums = torch.zeros(max(slabs, 1) * 2, dtype=torch.float64, device=DEV)
    d = GemmDesc()
    d.M, d.N, d.K, d.Cin, d.taps = M, N, K, K, 1
    d.I1, d.I0, d.m1, d.m0, d.J1 = 1, (stats_rows or M), 1, 1, 1
    d.J0 = d.I0
    d.xs_b, d.xs_1, d.xs_0, d.xs_c = d.I0 * lda, 0, lda, 1
    d.os_b, d.os_1, d.os_0 = d.I0 * n_out, 0, n_out
    d.x, d.w, d.bias, d.out = ptr(xfull), ptr(w), ptr(b), ptr(out)
    d.act, d.resid, d.scale = act, ptr(r), ptr(sc)
    d.stats_out = ptr(sums) if stats_rows else None
    d.stat_div, d.stat_mul, d.stat_mod = (stats_rows or M), 1, 1
    d.math = math_mode
    if math_mode in (_lib.MATH_BF16X3, _lib.MATH_BF16):     # pre-split bf16 weight planes (engine.Engine._w16)
        w_hi = w.to(torch.bfloat16)
        w_lo = (w - w_hi.float()).to(torch.bfloat16)
        d.w16_hi, d.w16_lo = ptr(w_hi), ptr(w_lo)
    _lib.call("bd_conv_gemm", C.byref(d), 0)
    torch.cuda.synchronize()
    want = x.double() @ w.double().t()
    if bias:
        want = want + b.double()
    if act == _lib.ACT_GELU:
        want = F.gelu(want)
    elif act == _lib.ACT_GLU:
        want = want[:, 0::2] * torch.sigmoid(want[:, 1::2])
    if resid:
        want = r.double() + sc.double() * want
    return out, want, sums, slabs


SHAPES = [(2688, 512, 512), (1344 * 3, 1536, 512), (1000, 2048, 512), (777, 512, 2048), (4096, 96, 96),
          (128, 64, 32), (5000, 384, 100), (3000, 96, 48), (2000, 16, 144), (1500, 48, 1152)]


MATHS = [(_lib.MATH_FP32, 3e-6), (_lib.MATH_TF32, 1.5e-3), (_lib.MATH_TF32X3, 2e-5), (_lib.MATH_BF16X3, 2e-5),
         (_lib.MATH_BF16, 6e-3)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("math_mode,tol", MATHS)
def test_plain_gemm(M, N, K, math_mode, tol):
    out, want, _, _ = run_plain(M, N, K, math_mode)
    assert not torch.isnan(out).any()
    e = rel_l2(out.cpu(), want.cpu())
    print(f"gemm math={math_mode} {M}x{N}x{K}: rel-L2 {e:.2e}")
    assert e < tol


@pytest.mark.parametrize("math_mode,tol", MATHS)
def test_fused_epilogues(math_mode, tol):
    out, want, _, _ = run_plain(2688, 2048, 512, math_mode, act=_lib.ACT_GELU)
    assert rel_l2(out.cpu(), want.cpu()) < tol
    out, want, _, _ = run_plain(3000, 768, 384, math_mode, act=_lib.ACT_GLU)
    assert rel_l2(out.cpu(), want.cpu()) < tol
    out, want, sums, slabs = run_plain(4 * 1344, 512, 2048, math_mode, resid=True, stats_rows=1344)
    assert rel_l2(out.cpu(), want.cpu()) < tol
    got = sums.view(slabs, 2).cpu()
    o = out.double().view(slabs, 1344, -1).cpu()
    assert torch.allclose(got[:, 0], o.sum(dim=(1, 2)), rtol=1e-6, atol=1e-4)
    assert torch.allclose(got[:, 1], (o ** 2).sum(dim=(1, 2)), rtol=1e-6)
    # strided A (a slice of a wider buffer), as the packed QKV projection output is consumed
    out, want, _, _ = run_plain(2000, 512, 512, math_mode, lda=1536)
    assert rel_l2(out.cpu(), want.cpu()) < tol


def test_tf32_arm_really_ran():
    """The TF32 result must differ from the fp32 one (else the dispatcher silently fell through)."""
    a, want, _, _ = run_plain(2688, 512, 512, _lib.MATH_FP32)
    b, _, _, _ = run_plain(2688, 512, 512, _lib.MATH_TF32)
    e = rel_l2(b.cpu(), a.cpu())
    assert 1e-5 < e < 1.5e-3
    c, _, _, _ = run_plain(2688, 512, 512, _lib.MATH_BF16)
    assert 2e-4 < rel_l2(c.cpu(), a.cpu()) < 6e-3
    d_, _, _, _ = run_plain(2688, 512, 512, _lib.MATH_BF16X3)
    assert 1e-7 < rel_l2(d_.cpu(), a.cpu()) < 2e-5


@pytest.mark.parametrize("math_mode,tol", [(_lib.MATH_TF32, 2e-3), (_lib.MATH_BF16X3, 5e-6)])
@pytest.mark.parametrize("Fr,dil", [(1, 1), (1, 2), (8, 1), (32, 2)])
def test_dconv_conv3_mma(Fr, dil, math_mode, tol):
    """Dedicated narrow DConv conv3 (mma.sync tf32 fragments) against an fp64 conv along the position axis,
    including the per-slab statistics; ragged row count (not a multiple of 32)."""
    B, T, C, hid = 3, 37, 48, 6
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, T, Fr, C, generator=g)
    w = torch.randn(hid, C, 3, generator=g) / (3 * C) ** 0.5
    b = torch.randn(hid, generator=g)
    w1 = w.permute(0, 2, 1).reshape(hid, 3 * C).contiguous()          # tap-major
    M = B * T * Fr
    xd, w1d, bd = x.to(DEV), w1.to(DEV), b.to(DEV)
    h = torch.full((M, 8), float("nan"), device=DEV)
    sums = torch.zeros(B * Fr, 2, dtype=torch.float64, device=DEV)
    _lib.call("bd_dconv_conv3", xd.data_ptr(), w1d.data_ptr(), bd.data_ptr(), h.data_ptr(), 8, sums.data_ptr(), M, C, hid,
              T * Fr, Fr, dil, math_mode, 0)
    torch.cuda.synchronize()
    xin = x.double().permute(0, 2, 3, 1).reshape(B * Fr, C, T)           # the reference's [(b f), c, t] view
    want = torch.nn.functional.conv1d(xin, w.double(), b.double(), padding=dil, dilation=dil)   # [(b f), hid, T]
    want_rows = want.reshape(B, Fr, hid, T).permute(0, 3, 1, 2).reshape(M, hid)
    got = h.cpu()
    assert torch.all(got[:, hid:] == 0)
    assert rel_l2(got[:, :hid], want_rows.float()) < tol               # single-pass tf32 / three-pass split
    s = want.reshape(B * Fr, -1)
    assert torch.allclose(sums[:, 0].cpu(), s.sum(1), rtol=0, atol=0.1)
    assert torch.allclose(sums[:, 1].cpu(), (s ** 2).sum(1), rtol=5e-3)


@pytest.mark.parametrize("math_mode,tol", [(_lib.MATH_TF32, 2e-3), (_lib.MATH_BF16X3, 5e-6)])
@pytest.mark.parametrize("channel_major", [0, 1])
def test_encoder_conv0_mma(channel_major, math_mode, tol):
    """First encoder layer (k=8, s=4, p=2, normalisation folded in, GELU) on mma.sync fragments vs torch fp64."""
    g = torch.Generator().manual_seed(9)
    B, cout = 3, 48
    if channel_major:
        cin, I1, Jin = 2, 1, 1003
        x = torch.randn(B, cin, Jin, generator=g)
        xin = x.double()                                                    # [B, cin, Jin]
    else:
        cin, I1, Jin = 4, 5, 64
        x = torch.randn(B, I1, Jin, cin, generator=g)
        xin = x.double().permute(0, 1, 3, 2).reshape(B * I1, cin, Jin)
    Io = (Jin + 3) // 4
    w = torch.randn(cout, cin, 8, generator=g) / (8 * cin) ** 0.5
    b = torch.randn(cout, generator=g)
    norm = torch.zeros(B, 8)
    norm[:, 0] = torch.randn(B, generator=g) * 0.1
    norm[:, 2] = 1.0 + 0.2 * torch.rand(B, generator=g)
    wp = w.permute(0, 2, 1).reshape(cout, 8 * cin).contiguous()
    out = torch.full((B, I1, Io, cout), float("nan"), device=DEV)
    xd, nd, wd, bd = x.to(DEV), norm.to(DEV), wp.to(DEV), b.to(DEV)
    _lib.call("bd_encoder_conv0", xd.data_ptr(), channel_major, nd.data_ptr(), 8, wd.data_ptr(), bd.data_ptr(),
              out.data_ptr(), B, I1, Io, Jin, cin, cout, math_mode, 0)
    torch.cuda.synchronize()
    rep = I1 if not channel_major else 1
    mean = norm[:, 0].double().repeat_interleave(rep).view(-1, 1, 1)
    rstd = norm[:, 2].double().repeat_interleave(rep).view(-1, 1, 1)
    xn = F.pad((xin - mean) * rstd, (2, 4 * Io + 4 - Jin + 2))              # zero pad AFTER normalisation
    want = F.gelu(F.conv1d(xn, w.double(), b.double(), stride=4))[..., :Io]  # [B*I1, cout, Io]
    want = want.reshape(B, I1, cout, Io).permute(0, 1, 3, 2)
    assert rel_l2(out.cpu(), want.float()) < tol


@pytest.mark.parametrize("M,N,K,act", [(2688, 1536, 512, _lib.ACT_NONE), (4000, 2048, 512, _lib.ACT_GELU),
                                       (1344 * 2, 512, 2048, _lib.ACT_NONE), (300, 64, 64, _lib.ACT_NONE)])
@pytest.mark.parametrize("out16", [0, 1])
def test_gemm_bf16_tensors(M, N, K, act, out16):
    """bf16 A tensor read by TMA in operand form (BD_MATH_BF16 + x_bf16), fp32 or bf16 output, fp32 residual: against an
    fp64 product of the SAME bf16-rounded operands (so only accumulation order and the output rounding differ)."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(M, K, generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    w_hi = w.to(torch.bfloat16)
    b = torch.randn(N, generator=g).to(DEV)
    r = torch.randn(M, N, generator=g).to(DEV) if not out16 else None
    sc = torch.randn(N, generator=g).to(DEV) if not out16 else None
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16 if out16 else torch.float32)
    d = GemmDesc()
    d.M, d.N, d.K, d.Cin, d.taps = M, N, K, K, 1
    d.I1, d.I0, d.m1, d.m0, d.J1, d.J0 = 1, M, 1, 1, 1, M
    d.xs_b, d.xs_1, d.xs_0, d.xs_c = M * K, 0, K, 1
    d.os_b, d.os_1, d.os_0 = M * N, 0, N
    d.x, d.w, d.bias, d.out = ptr(x), ptr(w), ptr(b), ptr(out)
    d.act, d.resid, d.scale = act, ptr(r), ptr(sc)
    d.stat_div, d.stat_mul, d.stat_mod = M, 1, 1
    d.math, d.w16_hi, d.x_bf16, d.out_bf16 = _lib.MATH_BF16, ptr(w_hi), 1, out16
    assert _lib.lib().bd_conv_gemm_arm(C.byref(d)) // 1000 == 64
    _lib.call("bd_conv_gemm", C.byref(d), 0)
    torch.cuda.synchronize()
    want = x.double() @ w_hi.double().t() + b.double()
    if act == _lib.ACT_GELU:
        want = F.gelu(want)
    if r is not None:
        want = r.double() + sc.double() * want
    e = rel_l2(out.float().cpu(), want.cpu())
    print(f"bf16-tensor gemm {M}x{N}x{K} out16={out16}: rel-L2 {e:.2e}")
    assert e < (3e-3 if out16 else 2e-6)


def test_layer_norm_bf16_output():
    g = torch.Generator().manual_seed(2)
    M, Cc = 3001, 512
    x = torch.randn(M, Cc, generator=g).to(DEV)
    gam, bet = torch.randn(Cc, generator=g).to(DEV), torch.randn(Cc, generator=g).to(DEV)
    pos = torch.randn(7, Cc, generator=g).to(DEV)
    y = torch.empty(M, Cc, device=DEV, dtype=torch.bfloat16)
    _lib.call("bd_layer_norm", ptr(x), ptr(y), ptr(gam), ptr(bet), ptr(pos), 7, M, Cc, 1, 0)
    torch.cuda.synchronize()
    want = F.layer_norm(x.double(), (Cc,), gam.double(), bet.double(), 1e-5) + pos.double()[torch.arange(M, device=DEV) % 7]
    assert rel_l2(y.float().cpu(), want.cpu()) < 3e-3
    assert (y.float() - want.float().to(torch.bfloat16).float()).abs().max() <= 1.6e-2 * want.abs().max()

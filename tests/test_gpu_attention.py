"""GPU: fused attention kernels (fp32 CUDA-core arm, tcgen05 TF32 and error-compensated 3xTF32 arms) against an fp64 softmax(QK^T/8)V."""
import pytest
import torch

from _fixtures import rel_l2
from demucs_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run(B, H, Tq, Tk, math_mode, packed=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    D = 64 * H
    if packed and Tq == Tk:     # self attention: one [B, T, 3D] projection buffer
        qkv = torch.randn(B, Tq, 3 * D, generator=g).to(DEV)
        q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
        ptrs = (qkv.data_ptr(), qkv.data_ptr() + 4 * D, qkv.data_ptr() + 8 * D)
        lds = (3 * D, 3 * D, 3 * D)
    else:                        # cross attention: q [B,Tq,D], packed kv [B,Tk,2D]
        qb = torch.randn(B, Tq, D, generator=g).to(DEV)
        kv = torch.randn(B, Tk, 2 * D, generator=g).to(DEV)
        q, k, v = qb, kv[..., :D], kv[..., D:]
        ptrs = (qb.data_ptr(), kv.data_ptr(), kv.data_ptr() + 4 * D)
        lds = (D, 2 * D, 2 * D)
    out = torch.full((B, Tq, D), float("nan"), device=DEV)
    ws = torch.empty(max(1, _lib.call_value("bd_attention_workspace", B, H, Tq, Tk, math_mode)), device=DEV)
    _lib.call("bd_attention", ptrs[0], ptrs[1], ptrs[2], out.data_ptr(), B, H, Tq, Tk, lds[0], lds[1], lds[2], D,
              math_mode, ws.data_ptr(), 0)
    torch.cuda.synchronize()
    qh = q.double().view(B, Tq, H, 64).transpose(1, 2)
    kh = k.double().view(B, Tk, H, 64).transpose(1, 2)
    vh = v.double().view(B, Tk, H, 64).transpose(1, 2)
    want = (torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, dim=-1) @ vh).transpose(1, 2).reshape(B, Tq, D)
    return out, want


SHAPES = [(2, 8, 2688, 2688), (2, 8, 1344, 1344), (1, 8, 2688, 1344), (1, 8, 1344, 2688), (3, 2, 100, 300),
          (2, 2, 352, 173), (1, 1, 128, 128), (1, 1, 129, 1)]


@pytest.mark.parametrize("B,H,Tq,Tk", SHAPES)
@pytest.mark.parametrize("math_mode,tol", [(_lib.MATH_FP32, 3e-6), (_lib.MATH_TF32, 2e-3),
                                           (_lib.MATH_TF32X3, 3e-5), (_lib.MATH_BF16X3, 3e-5),
                                           (_lib.MATH_BF16, 8e-3)])
def test_attention(B, H, Tq, Tk, math_mode, tol):
    out, want = run(B, H, Tq, Tk, math_mode)
    assert not torch.isnan(out).any()
    e = rel_l2(out.cpu(), want.cpu())
    print(f"attention math={math_mode} B{B} H{H} {Tq}x{Tk}: rel-L2 {e:.2e}")
    assert e < tol


@pytest.mark.parametrize("B,H,Tq,Tk", [(2, 8, 2688, 2688), (1, 8, 1344, 2688), (2, 2, 352, 173), (1, 1, 129, 1)])
def test_attention_bf16_tensors(B, H, Tq, Tk):
    """bf16 q / k / v read by TMA without a conversion pass, bf16 output: against fp64 on the same bf16 inputs."""
    g = torch.Generator().manual_seed(4)
    D = 64 * H
    if Tq == Tk:
        qkv = torch.randn(B, Tq, 3 * D, generator=g).to(DEV).to(torch.bfloat16)
        q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
        ptrs, lds = (qkv.data_ptr(), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D), (3 * D,) * 3
    else:
        qb = torch.randn(B, Tq, D, generator=g).to(DEV).to(torch.bfloat16)
        kv = torch.randn(B, Tk, 2 * D, generator=g).to(DEV).to(torch.bfloat16)
        q, k, v = qb, kv[..., :D], kv[..., D:]
        ptrs, lds = (qb.data_ptr(), kv.data_ptr(), kv.data_ptr() + 2 * D), (D, 2 * D, 2 * D)
    out = torch.full((B, Tq, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    _lib.call("bd_attention_bf16", ptrs[0], ptrs[1], ptrs[2], out.data_ptr(), B, H, Tq, Tk, lds[0], lds[1], lds[2], D, 0)
    torch.cuda.synchronize()
    qh = q.double().view(B, Tq, H, 64).transpose(1, 2)
    kh = k.double().view(B, Tk, H, 64).transpose(1, 2)
    vh = v.double().view(B, Tk, H, 64).transpose(1, 2)
    want = (torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, dim=-1) @ vh).transpose(1, 2).reshape(B, Tq, D)
    assert not torch.isnan(out.float()).any()
    e = rel_l2(out.float().cpu(), want.cpu())
    print(f"bf16-tensor attention B{B} H{H} {Tq}x{Tk}: rel-L2 {e:.2e}")
    assert e < 6e-3

// Epilogue shared by the CUDA-core and the tcgen05 implicit-GEMM kernels: everything that happens
// to an accumulator between the contraction and HBM (see bd_gemm_desc in include/demucs_b200.h).
// Two entry points: bd_epi_apply4 finishes 4 consecutive columns with 128-bit accesses (the hot
// path: every full-size layer qualifies), bd_epi_apply is the scalar form for odd geometries.
#pragma once
#include "common.cuh"
#include "../../include/demucs_b200.h"

struct EpiRow {
  long long obase;   // b*os_b + i1*os_1
  int i0;
  int rb_row;        // m % rowbias_period
  float e_mean, e_rstd;
};

__device__ __forceinline__ int bd_stat_slab(const bd_gemm_desc& d, long long m64) {
  const unsigned m = (unsigned)m64;
  return (int)((m / (unsigned)d.stat_div) * (unsigned)d.stat_mul + (m % (unsigned)d.stat_mod));
}

__device__ __forceinline__ EpiRow bd_epi_row(const bd_gemm_desc& d, long long m64) {
  EpiRow r;
  const unsigned m = (unsigned)m64;            // M < 2^31: 32-bit divisions
  const unsigned t = m / (unsigned)d.I0;
  r.i0 = (int)(m - t * (unsigned)d.I0);
  const unsigned b = t / (unsigned)d.I1;
  const int i1 = (int)(t - b * (unsigned)d.I1);
  r.obase = (long long)b * d.os_b + (long long)i1 * d.os_1;
  r.rb_row = d.rowbias ? (int)(m % (unsigned)d.rowbias_period) : 0;
  r.e_mean = 0.f;
  r.e_rstd = 1.f;
  if (d.e_stats) {
    const int sl = bd_stat_slab(d, m64);
    r.e_mean = __ldg(d.e_stats + 2 * (size_t)sl);
    r.e_rstd = __ldg(d.e_stats + 2 * (size_t)sl + 1);
  }
  return r;
}

// Can every epilogue operand be accessed as aligned float4 (float2 after GLU)?  Uniform per launch.
__device__ __forceinline__ bool bd_epi_vec_ok(const bd_gemm_desc& d) {
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  bool ok = d.N % 4 == 0 && d.os_0 % 4 == 0 && d.os_1 % 4 == 0 && d.os_b % 4 == 0;
  ok = ok && al16(d.out) && al16(d.bias) && al16(d.rowbias) && al16(d.resid) && al16(d.scale) && al16(d.addend) &&
       al16(d.e_gamma) && al16(d.e_beta);
  if (d.convt) ok = ok && ((d.N >> 2) % 4 == 0);
  if (d.act != BD_ACT_GLU && d.rowbias) ok = ok && (d.N % 4 == 0);
  return ok;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

// Finish accumulator columns n..n+3 (n % 4 == 0, n + 3 < N) of row r.  Adds the stored values to (s, q).
__device__ __forceinline__ void bd_epi_apply4(const bd_gemm_desc& d, const EpiRow& r, int n, float4 v, float& s,
                                              float& q) {
  if (d.bias) {
    const float4 b = ldg4(d.bias + n);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  if (d.e_stats) {
    const float4 g = ldg4(d.e_gamma + n), be = ldg4(d.e_beta + n);
    v.x = fmaf((v.x - r.e_mean) * r.e_rstd, g.x, be.x);
    v.y = fmaf((v.y - r.e_mean) * r.e_rstd, g.y, be.y);
    v.z = fmaf((v.z - r.e_mean) * r.e_rstd, g.z, be.z);
    v.w = fmaf((v.w - r.e_mean) * r.e_rstd, g.w, be.w);
  }
  if (d.act == BD_ACT_GLU) {
    float2 o2 = make_float2(v.x * bd_sigmoid(v.y), v.z * bd_sigmoid(v.w));
    const int no = n >> 1, Nout = d.N >> 1;
    const long long o = r.obase + (long long)r.i0 * d.os_0 + no;
    if (d.rowbias) {
      const float2 rb = ldg2(d.rowbias + (size_t)r.rb_row * Nout + no);
      o2.x += rb.x; o2.y += rb.y;
    }
    if (d.resid) {
      const float2 rs = ldg2(d.resid + o);
      float2 sc = make_float2(1.f, 1.f);
      if (d.scale) sc = ldg2(d.scale + no);
      o2.x = fmaf(sc.x, o2.x, rs.x); o2.y = fmaf(sc.y, o2.y, rs.y);
    }
    if (d.addend) {
      const float2 ad = ldg2(d.addend + o);
      o2.x += ad.x; o2.y += ad.y;
    }
    if (d.out) *reinterpret_cast<float2*>(d.out + o) = o2;
    s += o2.x + o2.y;
    q = fmaf(o2.x, o2.x, fmaf(o2.y, o2.y, q));
    return;
  }
  if (d.act == BD_ACT_GELU) {
    v.x = bd_gelu(v.x); v.y = bd_gelu(v.y); v.z = bd_gelu(v.z); v.w = bd_gelu(v.w);
  }
  int no = n, Nout = d.N;
  long long o;
  if (d.convt) {
    const int Cout = d.N >> 2;
    const int rr = n / Cout;
    const int o0 = 4 * r.i0 + rr - (d.convt == 1 ? 2 : 0);
    if (o0 < 0 || o0 >= d.O0) return;
    no = n - rr * Cout;
    Nout = Cout;
    o = r.obase + (long long)o0 * d.os_0 + no;
  } else {
    o = r.obase + (long long)r.i0 * d.os_0 + no;
  }
  if (d.rowbias) {
    const float4 rb = ldg4(d.rowbias + (size_t)r.rb_row * Nout + no);
    v.x += rb.x; v.y += rb.y; v.z += rb.z; v.w += rb.w;
  }
  if (d.resid) {
    const float4 rs = ldg4(d.resid + o);
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (d.scale) sc = ldg4(d.scale + no);
    v.x = fmaf(sc.x, v.x, rs.x); v.y = fmaf(sc.y, v.y, rs.y); v.z = fmaf(sc.z, v.z, rs.z); v.w = fmaf(sc.w, v.w, rs.w);
  }
  if (d.addend) {
    const float4 ad = ldg4(d.addend + o);
    v.x += ad.x; v.y += ad.y; v.z += ad.z; v.w += ad.w;
  }
  if (d.out) *reinterpret_cast<float4*>(d.out + o) = v;
  s += (v.x + v.y) + (v.z + v.w);
  q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, q))));
}

// Scalar form: finish accumulator `acc` of column n (and `acc_gate` of column n+1 for GLU, n even).
// Returns false when nothing is stored (GLU gate column / cropped transposed-conv position).
__device__ __forceinline__ bool bd_epi_apply(const bd_gemm_desc& d, const EpiRow& r, int n, float acc, float acc_gate,
                                             float& stored) {
  float v = acc + (d.bias ? __ldg(d.bias + n) : 0.f);
  if (d.e_stats) v = fmaf((v - r.e_mean) * r.e_rstd, __ldg(d.e_gamma + n), __ldg(d.e_beta + n));
  int no = n;
  int Nout = d.N;
  if (d.act == BD_ACT_GLU) {
    if (n & 1) return false;
    float g = acc_gate + (d.bias ? __ldg(d.bias + n + 1) : 0.f);
    if (d.e_stats) g = fmaf((g - r.e_mean) * r.e_rstd, __ldg(d.e_gamma + n + 1), __ldg(d.e_beta + n + 1));
    v = v * bd_sigmoid(g);
    no = n >> 1;
    Nout = d.N >> 1;
  } else if (d.act == BD_ACT_GELU) {
    v = bd_gelu(v);
  }
  long long o;
  if (d.convt) {
    const int Cout = d.N >> 2;
    const int rr = n / Cout;
    const int o0 = 4 * r.i0 + rr - (d.convt == 1 ? 2 : 0);
    if (o0 < 0 || o0 >= d.O0) return false;
    no = n - rr * Cout;
    Nout = Cout;
    o = r.obase + (long long)o0 * d.os_0 + no;
  } else {
    o = r.obase + (long long)r.i0 * d.os_0 + no;
  }
  if (d.rowbias) v += __ldg(d.rowbias + (size_t)r.rb_row * Nout + no);
  if (d.resid) v = fmaf(d.scale ? __ldg(d.scale + no) : 1.f, v, __ldg(d.resid + o));
  if (d.addend) v += __ldg(d.addend + o);
  if (d.out) d.out[o] = v;
  stored = v;
  return true;
}

// Epilogue shared by the CUDA-core and the tcgen05 implicit-GEMM kernels: everything that happens
// to an accumulator between the contraction and HBM (see bd_gemm_desc in include/demucs_b200.h).
#pragma once
#include "common.cuh"
#include "../../include/demucs_b200.h"

struct EpiRow {
  long long obase;   // b*os_b + i1*os_1
  int i0;
  int rb_row;        // m % rowbias_period
};

__device__ __forceinline__ EpiRow bd_epi_row(const bd_gemm_desc& d, long long m64) {
  EpiRow r;
  const unsigned m = (unsigned)m64;            // M < 2^31: 32-bit divisions
  const unsigned t = m / (unsigned)d.I0;
  r.i0 = (int)(m - t * (unsigned)d.I0);
  const unsigned b = t / (unsigned)d.I1;
  const int i1 = (int)(t - b * (unsigned)d.I1);
  r.obase = (long long)b * d.os_b + (long long)i1 * d.os_1;
  r.rb_row = d.rowbias ? (int)(m % (unsigned)d.rowbias_period) : 0;
  return r;
}

__device__ __forceinline__ int bd_stat_slab(const bd_gemm_desc& d, long long m64) {
  const unsigned m = (unsigned)m64;
  return (int)((m / (unsigned)d.stat_div) * (unsigned)d.stat_mul + (m % (unsigned)d.stat_mod));
}

// Finish accumulator `acc` of column n (and `acc_gate` of column n+1 for GLU, n even) of row `r`.
// Returns false when nothing is stored (GLU gate column / cropped transposed-conv position).
__device__ __forceinline__ bool bd_epi_apply(const bd_gemm_desc& d, const EpiRow& r, int n, float acc, float acc_gate,
                                             float& stored) {
  float v = acc + (d.bias ? __ldg(d.bias + n) : 0.f);
  int no = n;
  int Nout = d.N;
  if (d.act == BD_ACT_GLU) {
    if (n & 1) return false;
    const float g = acc_gate + (d.bias ? __ldg(d.bias + n + 1) : 0.f);
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
  d.out[o] = v;
  stored = v;
  return true;
}

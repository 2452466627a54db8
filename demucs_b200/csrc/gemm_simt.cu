// K3/K4/K7 (exact-fp32 arm): implicit-GEMM convolution / linear kernel on the CUDA cores.
//
// One kernel covers every contraction of the HTDemucs forward (see bd_gemm_desc in
// include/demucs_b200.h): the im2col gather, zero padding, A-side GroupNorm+GELU or item
// normalisation, bias, GELU/GLU, frequency embedding, LayerScale+residual, skip add, the
// transposed-conv scatter+crop and the GroupNorm statistics of the result are all fused, so each
// layer reads its input once and writes its output once.
//
// This is the bit-faithful fp32 arm (FFMA, fp32 accumulate) used in "fp32" mode, for the
// HBM-bound layers (C_in <= 8, DConv hidden widths 6..48) whose K or N is too small to feed a
// tensor-core tile, and as the on-device cross-check of the tcgen05 arm (gemm_tc.cu).
#include "gemm_epilogue.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int NTHREADS = 256;

struct RowInfo {
  long long xbase;   // b * xs_b
  int j1, j0;        // i1*m1, i0*m0
  int b;
  int slab;
  bool valid;
};

template <int BN, int TN>
__global__ void __launch_bounds__(NTHREADS) conv_gemm_simt_kernel(const bd_gemm_desc d, int slab_len, int tiles_per_slab) {
  constexpr int TM = 8;
  constexpr int TX = BN / TN;          // threads along N
  static_assert(TX * (BM / TM) == NTHREADS, "thread layout");
  constexpr int BPT = BN * BK / NTHREADS;  // weight elements per thread per k-tile
  constexpr int APAD = 4, BPAD = 4;
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN + BPAD];
  __shared__ double red[64];

  const int tid = threadIdx.x;
  const int slab = blockIdx.x / tiles_per_slab;
  const int tile = blockIdx.x - slab * tiles_per_slab;
  const int n0 = blockIdx.y * BN;
  const long long m_tile = (long long)slab * slab_len + (long long)tile * BM;
  const int rows_left = slab_len - tile * BM;  // rows of this tile inside the slab

  // ---- per-thread A gather role: one row, 8 consecutive k -------------------------------------
  const int arow = tid & (BM - 1);
  const int akg = (tid >> 7) * 8;
  RowInfo ri;
  {
    long long m = m_tile + arow;
    ri.valid = arow < rows_left && m < d.M;
    const unsigned mm = ri.valid ? (unsigned)m : 0u;
    const unsigned t = mm / (unsigned)d.I0;
    const int i0 = (int)(mm - t * (unsigned)d.I0);
    ri.b = (int)(t / (unsigned)d.I1);
    const int i1 = (int)(t - (unsigned)ri.b * (unsigned)d.I1);
    ri.xbase = ri.b * d.xs_b;
    ri.j1 = i1 * d.m1;
    ri.j0 = i0 * d.m0;
    ri.slab = d.a_mode == BD_A_GN_GELU ? bd_stat_slab(d, mm) : 0;
  }
  float a_mean = 0.f, a_rstd = 1.f;
  if (d.a_mode == BD_A_GN_GELU && ri.valid) {  // slab map: see bd_gemm_desc
    a_mean = d.a_stats[2 * (size_t)ri.slab];
    a_rstd = d.a_stats[2 * (size_t)ri.slab + 1];
  } else if (d.a_mode == BD_A_ITEM_AFFINE && ri.valid) {
    a_mean = d.a_stats[(size_t)ri.b * d.a_stats_stride];
    a_rstd = d.a_stats[(size_t)ri.b * d.a_stats_stride + 2];
  }
  const bool vec8 = (d.Cin % 8 == 0) && d.xs_c == 1;
  const bool vec4 = (d.Cin % 4 == 0) && d.xs_c == 1;

  auto tap_ptr = [&](int tap, bool& ok) -> const float* {
    int j1 = ri.j1 + d.d1[tap], j0 = ri.j0 + d.d0[tap];
    ok = ri.valid && j1 >= 0 && j1 < d.J1 && j0 >= 0 && j0 < d.J0;
    return d.x + ri.xbase + (long long)j1 * d.xs_1 + (long long)j0 * d.xs_0;
  };
  auto xform = [&](float v, int ci) -> float {
    if (d.a_mode == BD_A_GN_GELU) return bd_gelu(fmaf((v - a_mean) * a_rstd, __ldg(d.a_gamma + ci), __ldg(d.a_beta + ci)));
    if (d.a_mode == BD_A_ITEM_AFFINE) return (v - a_mean) * a_rstd;
    return v;
  };
  auto load_a = [&](int k0, float* r) {
    const int k = k0 + akg;
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = 0.f;
    if (k >= d.K) return;
    if (vec8) {
      int tap = k / d.Cin, ci = k - tap * d.Cin;
      bool ok;
      const float* p = tap_ptr(tap, ok);
      if (ok) {
        float4 u0 = __ldg(reinterpret_cast<const float4*>(p + ci));
        float4 u1 = __ldg(reinterpret_cast<const float4*>(p + ci + 4));
        r[0] = u0.x; r[1] = u0.y; r[2] = u0.z; r[3] = u0.w;
        r[4] = u1.x; r[5] = u1.y; r[6] = u1.z; r[7] = u1.w;
        if (d.a_mode != BD_A_NONE) {
#pragma unroll
          for (int i = 0; i < 8; ++i) r[i] = xform(r[i], ci + i);
        }
      }
    } else if (vec4) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int kk = k + 4 * h;
        if (kk >= d.K) break;
        int tap = kk / d.Cin, ci = kk - tap * d.Cin;
        bool ok;
        const float* p = tap_ptr(tap, ok);
        if (ok) {
          float4 u = __ldg(reinterpret_cast<const float4*>(p + ci));
          r[4 * h + 0] = u.x; r[4 * h + 1] = u.y; r[4 * h + 2] = u.z; r[4 * h + 3] = u.w;
          if (d.a_mode != BD_A_NONE) {
#pragma unroll
            for (int i = 0; i < 4; ++i) r[4 * h + i] = xform(r[4 * h + i], ci + i);
          }
        }
      }
    } else {
      int tap = k / d.Cin, ci = k - tap * d.Cin;
      bool ok;
      const float* p = tap_ptr(tap, ok);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (k + i < d.K) {
          if (ok) r[i] = xform(__ldg(p + (long long)ci * d.xs_c), ci);
          if (++ci == d.Cin) {
            ci = 0;
            ++tap;
            if (k + i + 1 < d.K) p = tap_ptr(tap, ok);
          }
        }
      }
    }
  };

  // ---- per-thread weight load role ---------------------------------------------------------------
  const int bn = tid % BN;
  const int bkq = (tid / BN) * BPT;
  const bool wvec = (d.K % 4 == 0) && (BPT % 4 == 0);
  auto load_b = [&](int k0, float* r) {
    const int n = n0 + bn, k = k0 + bkq;
#pragma unroll
    for (int i = 0; i < BPT; ++i) r[i] = 0.f;
    if (n >= d.N) return;
    const float* p = d.w + (size_t)n * d.K + k;
    if (wvec) {
#pragma unroll
      for (int i = 0; i < BPT; i += 4) {
        if (k + i < d.K) {
          float4 u = __ldg(reinterpret_cast<const float4*>(p + i));
          r[i] = u.x; r[i + 1] = u.y; r[i + 2] = u.z; r[i + 3] = u.w;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < BPT; ++i)
        if (k + i < d.K) r[i] = __ldg(p + i);
    }
  };

  // ---- compute role --------------------------------------------------------------------------------
  const int tx = tid % TX, ty = tid / TX;
  // rows: two groups of 4 separated by BM/2; cols: groups of min(TN,4) separated by BN/2 when TN == 8
  auto row_of = [&](int i) { return (i >> 2) * (BM / 2) + ty * 4 + (i & 3); };
  constexpr int CG = TN >= 4 ? 4 : TN;         // contiguous columns per group
  constexpr int NG = TN / CG;                  // column groups per thread
  auto col_of = [&](int j) { return (j / CG) * (BN / NG) + tx * CG + (j % CG); };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[8], rb[BPT];
  load_a(0, ra);
  load_b(0, rb);
#pragma unroll
  for (int i = 0; i < 8; ++i) As[0][akg + i][arow] = ra[i];
#pragma unroll
  for (int i = 0; i < BPT; ++i) Bs[0][bkq + i][bn] = rb[i];
  __syncthreads();

  const int nk = (d.K + BK - 1) / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load_a((kt + 1) * BK, ra);
      load_b((kt + 1) * BK, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float4 u = *reinterpret_cast<const float4*>(&As[cur][k][g * (BM / 2) + ty * 4]);
        a[4 * g] = u.x; a[4 * g + 1] = u.y; a[4 * g + 2] = u.z; a[4 * g + 3] = u.w;
      }
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const float* p = &Bs[cur][k][g * (BN / NG) + tx * CG];
        if constexpr (CG == 4) {
          float4 u = *reinterpret_cast<const float4*>(p);
          b[4 * g] = u.x; b[4 * g + 1] = u.y; b[4 * g + 2] = u.z; b[4 * g + 3] = u.w;
        } else if constexpr (CG == 2) {
          float2 u = *reinterpret_cast<const float2*>(p);
          b[0] = u.x; b[1] = u.y;
        } else {
          b[0] = *p;
        }
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      const int nxt = cur ^ 1;
#pragma unroll
      for (int i = 0; i < 8; ++i) As[nxt][akg + i][arow] = ra[i];
#pragma unroll
      for (int i = 0; i < BPT; ++i) Bs[nxt][bkq + i][bn] = rb[i];
    }
    __syncthreads();
  }

  // ---- epilogue ----------------------------------------------------------------------------------------
  // Statistics: tiles are slab-aligned when every row of a tile shares one GroupNorm slab (stat_mod == 1);
  // otherwise (frequency branch, slab = (b, fr) changes with every row) each row reduces across the TX
  // threads that share it and issues its own pair of atomics.
  const bool row_stats = d.stats_out && d.stat_mod != 1;
  const bool vec = bd_epi_vec_ok(d);
  EpiCol ecol[TN >= 4 ? TN / 4 : 1];
  if (TN >= 4 && vec) {
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      const int n = n0 + col_of(j);
      if (n < d.N) ecol[j / 4] = bd_epi_cols4(d, n);
    }
  }
  double ssum = 0.0, ssq = 0.0;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = row_of(i);
    const long long m = m_tile + r;
    const bool ok = r < rows_left && m < d.M;
    float rs = 0.f, rq = 0.f;
    if (ok) {
      const EpiRow er = bd_epi_row(d, m);
      if (TN >= 4 && vec) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const int n = n0 + col_of(j);
          if (n < d.N) bd_epi_apply4(d, er, ecol[j / 4], make_float4(acc[i][j], acc[i][(j + 1) % TN],
                                                                        acc[i][(j + 2) % TN], acc[i][(j + 3) % TN]), rs, rq);
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int n = n0 + col_of(j);
          if (n >= d.N) continue;
          float v;
          if (bd_epi_apply(d, er, n, acc[i][j], acc[i][(j + 1) % TN], v)) {
            rs += v;
            rq = fmaf(v, v, rq);
          }
        }
      }
    }
    if (row_stats) {   // the TX = 16 threads of a row are 16 consecutive lanes
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        rs += __shfl_xor_sync(0xffffffffu, rs, o);
        rq += __shfl_xor_sync(0xffffffffu, rq, o);
      }
      if (ok && tx == 0) {
        const int sl = bd_stat_slab(d, m);
        atomicAdd(&d.stats_out[2 * (size_t)sl], (double)rs);
        atomicAdd(&d.stats_out[2 * (size_t)sl + 1], (double)rq);
      }
    } else {
      ssum += rs;
      ssq += rq;
    }
  }
  if (d.stats_out && !row_stats) {
    bd_block_sum2(ssum, ssq, red);
    if (tid == 0) {
      const int sl = bd_stat_slab(d, m_tile);
      atomicAdd(&d.stats_out[2 * (size_t)sl], ssum);
      atomicAdd(&d.stats_out[2 * (size_t)sl + 1], ssq);
    }
  }
}

template <int BN, int TN>
int launch(const bd_gemm_desc& d, int slab_len, long long slabs, cudaStream_t st) {
  int tiles_per_slab = (slab_len + BM - 1) / BM;
  long long gx = slabs * tiles_per_slab;
  int gy = (d.N + BN - 1) / BN;
  if (gx > 0x7fffffffLL || gy > 65535) {
    bd_set_error("bd_conv_gemm: grid too large (%lld x %d)", gx, gy);
    return BD_ERR_ARG;
  }
  conv_gemm_simt_kernel<BN, TN><<<dim3((unsigned)gx, gy), NTHREADS, 0, st>>>(d, slab_len, tiles_per_slab);
  return bd_check_launch("conv_gemm_simt_kernel");
}

}  // namespace

int bd_conv_gemm_simt(const bd_gemm_desc* dp, void* stream) {
  const bd_gemm_desc& d = *dp;
  BD_REQUIRE(d.M > 0 && d.N > 0 && d.K > 0, "bd_conv_gemm: empty problem M=%d N=%d K=%d", d.M, d.N, d.K);
  BD_REQUIRE(d.taps >= 1 && d.taps <= BD_MAX_TAPS && d.K == d.taps * d.Cin, "bd_conv_gemm: K != taps*Cin");
  BD_REQUIRE(d.I0 > 0 && d.I1 > 0 && d.M % ((long long)d.I0 * d.I1) == 0, "bd_conv_gemm: M not a multiple of I1*I0");
  BD_REQUIRE(d.x && d.w && (d.out || d.stats_out), "bd_conv_gemm: null tensor");
  BD_REQUIRE(!d.e_stats || (d.e_gamma && d.e_beta && d.stat_div > 0), "bd_conv_gemm: epilogue GroupNorm needs affine + slab map");
  BD_REQUIRE(d.act != BD_ACT_GLU || (d.N % 2 == 0 && !d.convt), "bd_conv_gemm: GLU needs even N and no convt");
  BD_REQUIRE(!d.convt || d.N % 4 == 0, "bd_conv_gemm: convt needs N = 4*Cout");
  BD_REQUIRE(d.a_mode == BD_A_NONE || d.a_stats, "bd_conv_gemm: a_mode without a_stats");
  BD_REQUIRE(d.a_mode != BD_A_GN_GELU || (d.a_gamma && d.a_beta && d.taps == 1), "bd_conv_gemm: GN prologue needs affine + 1 tap");
  BD_REQUIRE(!d.rowbias || d.rowbias_period > 0, "bd_conv_gemm: rowbias without period");
  cudaStream_t st = (cudaStream_t)stream;
  BD_REQUIRE((!d.stats_out && d.a_mode != BD_A_GN_GELU) || (d.stat_div > 0 && d.stat_mul > 0 && d.stat_mod > 0),
             "bd_conv_gemm: statistics requested without a slab map");
  BD_REQUIRE(!d.stats_out || d.stat_mod != 1 || d.M % d.stat_div == 0, "bd_conv_gemm: M not a multiple of stat_div");
  // tiles never straddle a statistics slab when a whole tile reduces into one slab
  const bool slabbed = d.stats_out != nullptr && d.stat_mod == 1;
  const int slab_len = slabbed ? d.stat_div : d.M;
  const long long slabs = slabbed ? d.M / d.stat_div : 1;
  const bool glu = d.act == BD_ACT_GLU;
  if (d.N > 64) return launch<128, 8>(d, slab_len, slabs, st);
  if (d.N > 32) return launch<64, 4>(d, slab_len, slabs, st);
  if (d.N > 16 || glu) return launch<32, 2>(d, slab_len, slabs, st);
  return launch<16, 1>(d, slab_len, slabs, st);
}

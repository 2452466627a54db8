// DConv expansion stage (reference demucs/demucs.py:138-142,151-153 + LayerScale transformer.py:236-255):
//     g = gelu(GroupNorm(h));  u = W2 g + b2  (hidden -> 2C, 1x1 conv);  x += scale * GLU(GroupNorm(u))
// The [.., 2C] tensor u is 16x wider than h and never touches HBM: pass 1 (statistics) and pass 2
// (update) both recompute it from the narrow h.  With K = C/8 in 6..48 the contraction is far too thin
// for a tensor-core tile (the epilogue is the whole kernel), so this is a CUDA-core kernel shaped by
// HBM traffic: one thread = one row, the row's g vector lives in registers, W2 (transposed, <=147 KB)
// is broadcast from shared memory as float4, and x is updated in place one float4 at a time.
//   traffic / row: pass 1 reads hid floats; pass 2 reads hid + C and writes C floats.
#include "common.cuh"
#include "../../include/demucs_b200.h"

namespace {

constexpr int MAX_HID = 48;

template <int HID_T, bool FINAL, int DC_THREADS>
__global__ void __launch_bounds__(DC_THREADS) dconv_expand_kernel(
    const float* __restrict__ h, int ldh, int hid_rt, const float* __restrict__ mr1, const float* __restrict__ g1,
    const float* __restrict__ be1, const float* __restrict__ w2t /*[hid][2C] interleaved*/, const float* __restrict__ b2,
    double* __restrict__ sums2, const float* __restrict__ mr2, const float* __restrict__ g2, const float* __restrict__ be2,
    const float* __restrict__ scale, float* __restrict__ x, long long M, int C, long long rows_per_item,
    int slabs_per_item, int chunks) {
  extern __shared__ __align__(16) float sm[];
  const int hid = HID_T > 0 ? HID_T : hid_rt;
  const int N = 2 * C;
  float* sW = sm;                         // [hid][N]
  float* sB = sW + hid * N;               // [N]
  float* sG = sB + N;                     // [N]   (FINAL)
  float* sBe = sG + N;                    // [N]   (FINAL)
  float* sS = sBe + N;                    // [C]   (FINAL)
  float* sT = sS + C + (threadIdx.x >> 5) * (32 * 33);   // [32 rows][33] per warp (FINAL)
  for (int i = threadIdx.x; i < hid * N; i += DC_THREADS) sW[i] = __ldg(w2t + i);
  for (int i = threadIdx.x; i < N; i += DC_THREADS) {
    sB[i] = __ldg(b2 + i);
    if (FINAL) {
      sG[i] = __ldg(g2 + i);
      sBe[i] = __ldg(be2 + i);
    }
  }
  if (FINAL)
    for (int i = threadIdx.x; i < C; i += DC_THREADS) sS[i] = __ldg(scale + i);
  __syncthreads();

  // work item = (row, channel chunk): deep layers have few rows but wide C, so a row is split into `chunks`
  // column ranges to keep enough threads in flight (each recomputes the row's short g vector).  Chunk-major
  // order with the row count padded to a multiple of 32: the lanes of a warp work on 32 consecutive rows of
  // the SAME column range, so the shared-memory weight reads are warp-uniform broadcasts.
  const int ncols = N / chunks;           // interleaved columns per chunk (multiple of 4)
  const long long Mp = (M + 31) & ~31LL;
  const long long items = Mp * chunks;
  long long run_slab = -1;                // running statistics (time branch: consecutive rows share a slab)
  float run_s = 0.f, run_q = 0.f;
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * DC_THREADS;
  for (long long it0 = (long long)blockIdx.x * DC_THREADS; it0 < items; it0 += stride) {
    const long long it = it0 + threadIdx.x;
    const int chunk = (int)(it / Mp);                            // warp-uniform
    const long long m = it - (long long)chunk * Mp;
    const bool live = it < items && m < M;
    const long long m_w = m - lane;                              // first row of this warp
    const int n_lo = chunk * ncols;
    const long long slab = live ? (m / rows_per_item) * slabs_per_item + (m % slabs_per_item) : 0;
    float g[HID_T > 0 ? HID_T : MAX_HID];
    float mean2 = 0.f, rstd2 = 1.f;
    if (live) {
      const float mean1 = __ldg(mr1 + 2 * slab), rstd1 = __ldg(mr1 + 2 * slab + 1);
      const float* hr = h + m * ldh;
#pragma unroll
      for (int k = 0; k < (HID_T > 0 ? HID_T : MAX_HID); ++k)
        if (k < hid) g[k] = bd_gelu(fmaf((__ldg(hr + k) - mean1) * rstd1, __ldg(g1 + k), __ldg(be1 + k)));
      if (FINAL) {
        mean2 = __ldg(mr2 + 2 * slab);
        rstd2 = __ldg(mr2 + 2 * slab + 1);
      }
    } else {
#pragma unroll
      for (int k = 0; k < (HID_T > 0 ? HID_T : MAX_HID); ++k) g[k] = 0.f;
    }
    float s = 0.f, q = 0.f;
    if (it0 + (threadIdx.x & ~31) < items) {                     // warp-uniform: this warp has work
      for (int n = n_lo; n < n_lo + ncols; n += 4) {             // 4 interleaved columns = (value, gate) x 2 channels
        float4 u = *reinterpret_cast<const float4*>(sB + n);
#pragma unroll
        for (int k = 0; k < (HID_T > 0 ? HID_T : MAX_HID); ++k) {
          if (k < hid) {
            const float4 w = *reinterpret_cast<const float4*>(sW + k * N + n);
            u.x = fmaf(w.x, g[k], u.x); u.y = fmaf(w.y, g[k], u.y);
            u.z = fmaf(w.z, g[k], u.z); u.w = fmaf(w.w, g[k], u.w);
          }
        }
        if (FINAL) {
          const float4 ga = *reinterpret_cast<const float4*>(sG + n), be = *reinterpret_cast<const float4*>(sBe + n);
          const float a0 = fmaf((u.x - mean2) * rstd2, ga.x, be.x), t0 = fmaf((u.y - mean2) * rstd2, ga.y, be.y);
          const float a1 = fmaf((u.z - mean2) * rstd2, ga.z, be.z), t1 = fmaf((u.w - mean2) * rstd2, ga.w, be.w);
          const float2 sc = *reinterpret_cast<const float2*>(sS + (n >> 1));
          // x is updated 32 channels at a time: every lane parks the increments of ITS row in the warp's staging
          // tile, then the warp walks its 32 rows with lane = channel -- each access to x is one 128-byte run
          const int cc = (n - n_lo) >> 1;                        // channel inside the chunk
          sT[lane * 33 + (cc & 31)] = sc.x * a0 * bd_sigmoid(t0);
          sT[lane * 33 + (cc & 31) + 1] = sc.y * a1 * bd_sigmoid(t1);
          if ((cc & 31) == 30 || n + 4 >= n_lo + ncols) {        // block of 32 channels complete (or chunk end)
            const int c_lo = cc & ~31, width = cc + 2 - c_lo;
            __syncwarp();
            if (lane < width) {
              float* xc = x + (n_lo >> 1) + c_lo + lane + m_w * C;
              const int nrows = (int)(M - m_w < 32 ? M - m_w : 32);
#pragma unroll 1
              for (int r0 = 0; r0 < 32; r0 += 8) {               // 8 row loads in flight before the first store
                float xv[8];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                  if (r0 + r < nrows) xv[r] = xc[(long long)(r0 + r) * C];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                  if (r0 + r < nrows) xc[(long long)(r0 + r) * C] = xv[r] + sT[(r0 + r) * 33 + lane];
              }
            }
            __syncwarp();
          }
        } else {
          s += (u.x + u.y) + (u.z + u.w);
          q = fmaf(u.x, u.x, fmaf(u.y, u.y, fmaf(u.z, u.z, fmaf(u.w, u.w, q))));
        }
      }
    }
    if (!FINAL) {
      // the whole warp usually sits in one slab (always, in the time branch): one atomic pair per warp and
      // slab change instead of one per row
      const long long slab0 = __shfl_sync(0xffffffffu, slab, 0);
      const bool uniform = __all_sync(0xffffffffu, !live || slab == slab0);
      if (uniform) {
        const float ws = bd_warp_sum(live ? s : 0.f), wq = bd_warp_sum(live ? q : 0.f);
        if (lane == 0 && live) {
          if (slab0 != run_slab) {
            if (run_slab >= 0) {
              atomicAdd(&sums2[2 * run_slab], (double)run_s);
              atomicAdd(&sums2[2 * run_slab + 1], (double)run_q);
            }
            run_slab = slab0;
            run_s = run_q = 0.f;
          }
          run_s += ws;
          run_q += wq;
        }
      } else if (live) {
        atomicAdd(&sums2[2 * slab], (double)s);
        atomicAdd(&sums2[2 * slab + 1], (double)q);
      }
    }
  }
  if (!FINAL && lane == 0 && run_slab >= 0) {
    atomicAdd(&sums2[2 * run_slab], (double)run_s);
    atomicAdd(&sums2[2 * run_slab + 1], (double)run_q);
  }
}

template <int HID_T, bool FINAL, int THREADS>
int launch_one(const float* h, int ldh, int hid, const float* mr1, const float* g1, const float* be1, const float* w2t,
               const float* b2, double* sums2, const float* mr2, const float* g2, const float* be2, const float* scale,
               float* x, long long M, int C, long long rpi, int spi, int ctas_per_sm, cudaStream_t st) {
  const int smem = (hid * 2 * C + 2 * C * 3 + C + (FINAL ? (THREADS / 32) * 32 * 33 : 0)) * (int)sizeof(float);
  // split rows into channel chunks until ~256k work items exist (chunk = multiple of 2 channels)
  int chunks = 1;
  while (M * chunks < 262144 && chunks < 16 && (C % (4 * chunks)) == 0 && C / (2 * chunks) >= 8) chunks *= 2;
  long long grid = ((((M + 31) & ~31LL) * chunks + THREADS - 1) / THREADS);
  const long long cap = 148LL * ctas_per_sm * 2;   // grid-stride: weights are staged once per CTA
  if (grid > cap) grid = cap;
  cudaError_t e = cudaFuncSetAttribute(dconv_expand_kernel<HID_T, FINAL, THREADS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    bd_set_error("bd_dconv_expand: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return BD_ERR_CUDA;
  }
  dconv_expand_kernel<HID_T, FINAL, THREADS><<<(unsigned)grid, THREADS, smem, st>>>(
      h, ldh, hid, mr1, g1, be1, w2t, b2, sums2, mr2, g2, be2, scale, x, M, C, rpi, spi, chunks);
  return bd_check_launch("dconv_expand_kernel");
}

template <bool FINAL>
int launch_expand(const float* h, int ldh, int hid, const float* mr1, const float* g1, const float* be1, const float* w2t,
                  const float* b2, double* sums2, const float* mr2, const float* g2, const float* be2, const float* scale,
                  float* x, long long M, int C, long long rpi, int spi, cudaStream_t st) {
#define BD_DC_ARGS h, ldh, hid, mr1, g1, be1, w2t, b2, sums2, mr2, g2, be2, scale, x, M, C, rpi, spi
  // narrow layers: many small CTAs; wide layers (W2 of 36..147 KB in shared memory): one big CTA per SM
  if (hid == 6) return launch_one<6, FINAL, 128>(BD_DC_ARGS, 8, st);
  if (hid == 12) return launch_one<12, FINAL, 128>(BD_DC_ARGS, 8, st);
  if (hid == 24) return launch_one<24, FINAL, 256>(BD_DC_ARGS, 2, st);
  if (hid == 48) return launch_one<48, FINAL, 512>(BD_DC_ARGS, 1, st);
  const int wbytes = hid * 2 * C * 4;
  if (wbytes > 64 * 1024) return launch_one<0, FINAL, 512>(BD_DC_ARGS, 1, st);
  return launch_one<0, FINAL, 128>(BD_DC_ARGS, wbytes > 24 * 1024 ? 2 : 8, st);
#undef BD_DC_ARGS
}


// ---- pass 1 through the Gram matrix ---------------------------------------------------------------------
// sum(u) and sum(u^2) over a slab are quadratic forms of the slab's g statistics:
//     sum_rows sum_n u_n   = sum_n (w_n . sg) + R sum_n b_n                    sg = sum_rows g
//     sum_rows sum_n u_n^2 = sum_n (w_n^T G w_n + 2 b_n (w_n . sg) + R b_n^2)   G  = sum_rows g g^T
// so pass 1 needs hid^2 products per row instead of hid * 2C (16x fewer) and never forms u at all.
// A warp walks rows r = (c*K + k)*P + 32*sb + lane of one item, P = max(32, slabs_per_item): the slab of a lane
// is the same in every iteration, lanes that share a slab (slabs_per_item < 32) are folded with shuffles, and
// each slab's owner lane adds its BLK x BLK block of G (and sg) to the fp64 workspace.
// The column sums of the quadratic forms depend on the weights only:
//     sum_n w_n^T G w_n = sum_ab G[a][b] WW[a][b],   WW = W2^T W2;   sum_n b_n (w_n . sg) = wb . sg;   ...
// wst = [WW (hid*hid) | ws = sum_n w_n (hid) | wb = sum_n b_n w_n (hid) | sum b_n | sum b_n^2], one warp per entry.
__device__ __forceinline__ void dconv_wstats(const float* __restrict__ w2t, const float* __restrict__ b2, int hid, int N,
                                             double* __restrict__ wst, int cta, int nctas) {
  const int lane = threadIdx.x & 31;
  const int E = hid * hid + 2 * hid + 2;
  for (int e = cta * 8 + (threadIdx.x >> 5); e < E; e += nctas * 8) {
    // operands of entry e: two rows of [w2t ; b2 ; 1]
    const float* px = nullptr;
    const float* py = nullptr;                                  // nullptr = the all-ones row
    if (e < hid * hid) {
      px = w2t + (size_t)(e / hid) * N;
      py = w2t + (size_t)(e % hid) * N;
    } else if (e < hid * hid + hid) {
      px = w2t + (size_t)(e - hid * hid) * N;
    } else if (e < hid * hid + 2 * hid) {
      px = w2t + (size_t)(e - hid * hid - hid) * N;
      py = b2;
    } else {
      px = b2;
      py = e == hid * hid + 2 * hid ? nullptr : b2;
    }
    double acc = 0.0;
    for (int n0 = 0; n0 < N; n0 += 256) {                       // 8 independent loads per lane in flight
      float x[8], y[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int n = n0 + 32 * u + lane;
        x[u] = n < N ? __ldg(px + n) : 0.f;
        y[u] = n < N ? (py ? __ldg(py + n) : 1.f) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc = fma((double)x[u], (double)y[u], acc);
    }
    acc = bd_warp_sum_d(acc);
    if (lane == 0) wst[e] = acc;
  }
}

template <int BLK>
__global__ void __launch_bounds__(256, (BLK <= 8 ? 2 : 1)) dconv_gram_kernel(const float* __restrict__ h, int ldh, int hid,
                                                         const float* __restrict__ mr1, const float* __restrict__ g1,
                                                         const float* __restrict__ be1, double* __restrict__ gram,
                                                         long long rpi, int spi, int items, int K, int ngroups,
                                                         int gram_ctas, const float* __restrict__ w2t,
                                                         const float* __restrict__ b2, int N, double* __restrict__ wst) {
  constexpr int NACC = BLK * BLK + BLK, LDA = NACC | 1;       // odd pitch: lanes hit different banks
  __shared__ float buf[32 * LDA];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((int)blockIdx.x >= gram_ctas) {
    // the trailing CTAs compute the weight-only column sums while the others stream h (see dconv_wstats below)
    dconv_wstats(w2t, b2, hid, N, wst, (int)blockIdx.x - gram_ctas, (int)gridDim.x - gram_ctas);
    return;
  }
  const int nblk = hid / BLK, nunits = nblk * (nblk + 1) / 2;
  const int P = spi > 32 ? spi : 32, nsb = P / 32;
  const int nown = spi < 32 ? spi : 32;                       // lanes that own a slab after the fold
  const long long nwork = (long long)items * nunits * nsb * ngroups;
  const int GS = hid * hid + hid;
  // CTA = 8 warps = 8 consecutive row chunks of one (item, block pair, slab block): their partial sums meet in
  // shared memory, so the fp64 workspace sees one atomic per value and CTA instead of one per warp
  for (long long wk = blockIdx.x; wk < nwork; wk += gram_ctas) {
    long long t = wk;
    const int c = (int)(t % ngroups) * 8 + warp; t /= ngroups;
    const int sb = (int)(t % nsb); t /= nsb;
    int unit = (int)(t % nunits);
    const int item = (int)(t / nunits);
    int bi = 0;
    while (unit >= nblk - bi) { unit -= nblk - bi; ++bi; }
    const int bj = bi + unit;
    const bool diag = bi == bj;
    for (int i = threadIdx.x; i < 32 * LDA; i += 256) buf[i] = 0.f;
    __syncthreads();
    float ga[BLK], bea[BLK], gb[BLK], beb[BLK];
#pragma unroll
    for (int a = 0; a < BLK; ++a) {
      ga[a] = __ldg(g1 + bi * BLK + a); bea[a] = __ldg(be1 + bi * BLK + a);
      gb[a] = __ldg(g1 + bj * BLK + a); beb[a] = __ldg(be1 + bj * BLK + a);
    }
    float acc[BLK][BLK], sg[BLK];
#pragma unroll
    for (int a = 0; a < BLK; ++a) {
      sg[a] = 0.f;
#pragma unroll
      for (int b = 0; b < BLK; ++b) acc[a][b] = 0.f;
    }
    const long long r_first = (long long)c * K * P + 32 * sb + lane;
    const int s_lane = (int)((32LL * sb + lane) % spi);
    const long long slab = (long long)item * spi + s_lane;
    const float mean1 = __ldg(mr1 + 2 * slab), rstd1 = __ldg(mr1 + 2 * slab + 1);
    // rows in batches of UB with every load issued before the first GELU (the chain load -> erf -> fma is long)
    constexpr int UB = BLK > 8 ? 2 : 4;
    for (int k0 = 0; k0 < K; k0 += UB) {
      if (r_first - lane + (long long)k0 * P >= rpi) break;     // warp-uniform
      float ra[UB][BLK], rb[UB][BLK];
      bool live[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const long long r = r_first + (long long)(k0 + u) * P;
        live[u] = k0 + u < K && r < rpi;
        const float* hr = h + ((long long)item * rpi + (live[u] ? r : 0)) * ldh;
        if (BLK % 4 == 0) {
#pragma unroll
          for (int a = 0; a < BLK; a += 4) {
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(hr + bi * BLK + a));
            ra[u][a] = q0.x; ra[u][a + 1] = q0.y; ra[u][a + 2] = q0.z; ra[u][a + 3] = q0.w;
            if (!diag) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(hr + bj * BLK + a));
              rb[u][a] = q.x; rb[u][a + 1] = q.y; rb[u][a + 2] = q.z; rb[u][a + 3] = q.w;
            }
          }
        } else {
#pragma unroll
          for (int a = 0; a < BLK; a += 2) {
            const float2 q0 = __ldg(reinterpret_cast<const float2*>(hr + bi * BLK + a));
            ra[u][a] = q0.x; ra[u][a + 1] = q0.y;
            if (!diag) {
              const float2 q = __ldg(reinterpret_cast<const float2*>(hr + bj * BLK + a));
              rb[u][a] = q.x; rb[u][a + 1] = q.y;
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if (!live[u]) continue;
        float va[BLK], vb[BLK];
#pragma unroll
        for (int a = 0; a < BLK; ++a) {
          va[a] = bd_gelu(fmaf((ra[u][a] - mean1) * rstd1, ga[a], bea[a]));
          sg[a] += va[a];
        }
        if (diag) {
#pragma unroll
          for (int a = 0; a < BLK; ++a) vb[a] = va[a];
        } else {
#pragma unroll
          for (int a = 0; a < BLK; ++a) vb[a] = bd_gelu(fmaf((rb[u][a] - mean1) * rstd1, gb[a], beb[a]));
        }
#pragma unroll
        for (int a = 0; a < BLK; ++a)
#pragma unroll
          for (int b = 0; b < BLK; ++b) acc[a][b] = fmaf(va[a], vb[b], acc[a][b]);
      }
    }
    // fold lanes that share a slab, then meet the other warps in shared memory
    for (int o = 16; o >= spi; o >>= 1) {
#pragma unroll
      for (int a = 0; a < BLK; ++a) {
        sg[a] += __shfl_xor_sync(0xffffffffu, sg[a], o);
#pragma unroll
        for (int b = 0; b < BLK; ++b) acc[a][b] += __shfl_xor_sync(0xffffffffu, acc[a][b], o);
      }
    }
    if (lane < nown && r_first - lane < rpi) {
#pragma unroll
      for (int a = 0; a < BLK; ++a) {
#pragma unroll
        for (int b = 0; b < BLK; ++b) atomicAdd(&buf[lane * LDA + a * BLK + b], acc[a][b]);
        atomicAdd(&buf[lane * LDA + BLK * BLK + a], sg[a]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nown * NACC; i += 256) {
      const int ln = i / NACC, idx = i - ln * NACC;
      const long long sl = (long long)item * spi + (32LL * sb + ln) % spi;
      double* G = gram + sl * GS;
      const float v = buf[ln * LDA + idx];
      if (idx < BLK * BLK) {
        const int a = idx / BLK, b2 = idx - a * BLK;
        atomicAdd(&G[(bi * BLK + a) * hid + bj * BLK + b2], (double)v);
      } else if (diag) {
        atomicAdd(&G[hid * hid + bi * BLK + idx - BLK * BLK], (double)v);
      }
    }
    __syncthreads();
  }
}

// sums2[slab] += (sum u, sum u^2) from the slab's Gram matrix: one CTA per slab.  The Gram entries are cleared
// after they are read, so the workspace is all zeros again when the call returns.
template <int BLK>
__global__ void __launch_bounds__(128) dconv_gram_eval_kernel(double* __restrict__ gram, const double* __restrict__ wst,
                                                              int hid, double* __restrict__ sums2, double R) {
  const long long slab = blockIdx.x;
  const int HH = hid * hid, GS = HH + hid;
  double* G = gram + slab * GS;
  double S = 0.0, Q = 0.0;
  for (int i = threadIdx.x; i < HH; i += 128) {
    const int a = i / hid, b = i - a * hid;
    const double g = (a / BLK <= b / BLK) ? G[i] : G[b * hid + a];   // only blocks with bi <= bj were accumulated
    Q = fma(g, wst[i], Q);
  }
  for (int a = threadIdx.x; a < hid; a += 128) {
    const double sg = G[HH + a];
    S = fma(wst[HH + a], sg, S);
    Q = fma(2.0 * wst[HH + hid + a], sg, Q);
  }
  __shared__ double red[2][4];
  S = bd_warp_sum_d(S);
  Q = bd_warp_sum_d(Q);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = S;
    red[1][threadIdx.x >> 5] = Q;
  }
  __syncthreads();                                             // every read of G is done
  for (int i = threadIdx.x; i < GS; i += 128) G[i] = 0.0;
  if (threadIdx.x == 0) {
    sums2[2 * slab] += red[0][0] + red[0][1] + red[0][2] + red[0][3] + R * wst[HH + 2 * hid];
    sums2[2 * slab + 1] += red[1][0] + red[1][1] + red[1][2] + red[1][3] + R * wst[HH + 2 * hid + 1];
  }
}

bool gram_supported(int hid, int ldh, long long rpi, int spi) {
  if (hid != 6 && hid != 12 && hid != 24 && hid != 48) return false;
  if (ldh % 2 != 0) return false;                              // float2 loads of the h rows
  if (spi <= 0 || rpi % spi != 0) return false;
  return spi >= 32 ? spi % 32 == 0 : 32 % spi == 0;
}

constexpr int GRAM_BLK = 6;

int launch_gram(const float* h, int ldh, int hid, const float* mr1, const float* g1, const float* be1, const float* w2t,
                const float* b2, double* sums2, double* gram, long long M, int C, long long rpi, int spi, cudaStream_t st) {
  constexpr int BLK = GRAM_BLK;
  const int items = (int)(M / rpi);
  const long long slabs = (long long)items * spi;
  const int GS = hid * hid + hid;
  double* wst = gram + slabs * GS;                            // weight-only sums live behind the Gram matrices
  const int P = spi > 32 ? spi : 32, nsb = P / 32, nblk = hid / BLK, nunits = nblk * (nblk + 1) / 2;
  const long long periods = (rpi + P - 1) / P;               // iterations needed to cover an item
  // enough CTAs (8 warps each) to fill the machine, at least 8 rows per lane to amortise the reduction
  const long long base = (long long)items * nunits * nsb;
  long long ngroups = (148LL * 4 + base - 1) / base;
  if (ngroups > (periods + 63) / 64) ngroups = (periods + 63) / 64;
  if (ngroups < 1) ngroups = 1;
  const int K = (int)((periods + 8 * ngroups - 1) / (8 * ngroups));
  ngroups = (periods + 8LL * K - 1) / (8LL * K);
  long long grid = base * ngroups;
  if (grid > 148LL * 8) grid = 148LL * 8;
  const int wst_ctas = 148;
  dconv_gram_kernel<BLK><<<(unsigned)grid + wst_ctas, 256, 0, st>>>(h, ldh, hid, mr1, g1, be1, gram, rpi, spi, items, K,
                                                                    (int)ngroups, (int)grid, w2t, b2, 2 * C, wst);
  if (bd_check_launch("dconv_gram_kernel") != BD_OK) return BD_ERR_CUDA;
  dconv_gram_eval_kernel<BLK><<<(unsigned)slabs, 128, 0, st>>>(gram, wst, hid, sums2, (double)(rpi / spi));
  return bd_check_launch("dconv_gram_eval_kernel");
}


// ---- pass 2 on the warp-level tensor cores (single-pass TF32 mode) ------------------------------------------------
// The 1x1 expansion is a [rows x (hid+1)] x [(hid+1) x 2C] product with hid = 6..48: one to seven UMMA_K steps,
// so on a tcgen05 tile the epilogue is the whole kernel.  mma.sync.m16n8k8.tf32 works on registers instead: the
// bias is folded in as row `hid` against a constant 1 in g, GroupNorm + GELU of h happen while the A fragments
// are built, a warp takes 32 rows (two 16-row fragments), and the accumulator fragment hands every thread adjacent
// (value, gate) column pairs of two rows -- a GLU pair -- so GroupNorm, GLU, LayerScale and the residual update
// happen in place on the fragment, and the results go through a per-warp staging tile so that x is updated with
// contiguous float4 runs.  ~2.4x fewer instructions per row than the FFMA kernel above, which stays in charge
// of the fp32 and 3xTF32 modes (it is exact, and faster than three MMA passes).
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// Error-compensated form ("3xTF32", the strict modes): x = hi + lo with hi = rn_tf32(x) and lo = x - hi (exact in fp32;
// the tensor core reads its top 10 mantissa bits), product = lo*hi + hi*lo + hi*hi, small terms first.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32x3_16x8x8(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                                  uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32_16x8x8(c, al, bh0, bh1);
  mma_tf32_16x8x8(c, ah, bl0, bl1);
  mma_tf32_16x8x8(c, ah, bh0, bh1);
}

// CTA = 8 warps that share one block of 96 interleaved columns (48 channels) of the expansion: its weights sit in
// shared memory as tf32 [96][8*KS + 4] (the pitch makes the B-fragment loads conflict-free), every warp walks
// 32-row tiles.  KS = k-steps of 8: hid 6 -> 1, 12 -> 2, 24 -> 4, 48 -> 7 (row `hid` carries the bias).
template <int HID, bool X3>
__global__ void __launch_bounds__(256) dconv_update_mma_kernel(
    const float* __restrict__ h, int ldh, const float* __restrict__ mr1, const float* __restrict__ g1,
    const float* __restrict__ be1, const float* __restrict__ w2t /*[hid][2C]*/, const float* __restrict__ b2,
    const float* __restrict__ mr2, const float* __restrict__ g2, const float* __restrict__ be2,
    const float* __restrict__ scale, float* __restrict__ x, long long M, int C, long long rows_per_item,
    int slabs_per_item) {
  constexpr int KS = (HID + 1 + 7) / 8, KP = 8 * KS + 4;
  constexpr int NB = 96, CB = 48, NT = NB / 8;     // columns / channels / column fragments per block
  constexpr int LDT = CB + 4;                      // staging pitch: the fragment scatter hits 32 different banks
  const int N = 2 * C;
  const int n_lo = blockIdx.y * NB, c_lo = blockIdx.y * CB;
  extern __shared__ __align__(16) float sm[];
  uint32_t* sW = reinterpret_cast<uint32_t*>(sm);  // [NB][KP] tf32 (X3: then the lo parts, [NB][KP])
  uint32_t* sWl = sW + NB * KP;
  float* sG = sm + (X3 ? 2 : 1) * NB * KP;         // [NB]
  float* sBe = sG + NB;                            // [NB]
  float* sS = sBe + NB;                            // [CB]
  float* sT = sS + CB + (threadIdx.x >> 5) * (32 * LDT);   // [32 rows][LDT] per warp
  for (int i = threadIdx.x; i < NB * 8 * KS; i += 256) {
    const int n = i / (8 * KS), k = i - n * (8 * KS);
    const float w = k < HID ? __ldg(w2t + (size_t)k * N + n_lo + n) : (k == HID ? __ldg(b2 + n_lo + n) : 0.f);
    if (X3) {
      split_tf32(w, sW[n * KP + k], sWl[n * KP + k]);
    } else {
      sW[n * KP + k] = to_tf32(w);
    }
  }
  for (int i = threadIdx.x; i < NB; i += 256) {
    sG[i] = __ldg(g2 + n_lo + i);
    sBe[i] = __ldg(be2 + n_lo + i);
  }
  for (int i = threadIdx.x; i < CB; i += 256) sS[i] = __ldg(scale + c_lo + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  float ga[2 * KS], bea[2 * KS];                   // GroupNorm affine of this thread's k indices: 8*ks + tig (+4)
#pragma unroll
  for (int j = 0; j < 2 * KS; ++j) {
    const int k = 8 * (j >> 1) + tig + 4 * (j & 1);
    ga[j] = k < HID ? __ldg(g1 + k) : 0.f;
    bea[j] = k < HID ? __ldg(be1 + k) : 0.f;
  }
  const long long ntiles = (M + 31) / 32;
  const long long wstride = (long long)gridDim.x * 8;
  for (long long tile = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); tile < ntiles; tile += wstride) {
    const long long m0 = tile * 32;
    // this thread's 4 rows: gid, gid + 8 (fragment 0), gid + 16, gid + 24 (fragment 1)
    uint32_t ah[2][KS][4], al[X3 ? 2 : 1][X3 ? KS : 1][4];
    float m2[4], r2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long m = m0 + gid + 8 * q;
      const long long mm = m < M ? m : m0;
      const long long slab = (mm / rows_per_item) * slabs_per_item + (mm % slabs_per_item);
      const float mean1 = __ldg(mr1 + 2 * slab), rstd1 = __ldg(mr1 + 2 * slab + 1);
      m2[q] = __ldg(mr2 + 2 * slab);
      r2[q] = __ldg(mr2 + 2 * slab + 1);
      const float* hr = h + mm * ldh;
      const int f = q >> 1, hi8 = q & 1;
      // fragment element order: a0 (row gid, k tig), a1 (row gid+8, k tig), a2 (row gid, k tig+4), a3 (row gid+8, k tig+4)
#pragma unroll
      for (int j = 0; j < 2 * KS; ++j) {
        const int k = 8 * (j >> 1) + tig + 4 * (j & 1);
        float v = k == HID ? 1.f : 0.f;
        if (k < HID) v = bd_gelu(fmaf((__ldg(hr + k) - mean1) * rstd1, ga[j], bea[j]));
        if (X3) {
          split_tf32(v, ah[f][j >> 1][2 * (j & 1) + hi8], al[X3 ? f : 0][X3 ? (j >> 1) : 0][2 * (j & 1) + hi8]);
        } else {
          ah[f][j >> 1][2 * (j & 1) + hi8] = to_tf32(v);
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int n = 8 * nt + 2 * tig;                              // (value, gate) columns of this thread
      const float2 gam = *reinterpret_cast<const float2*>(sG + n), bet = *reinterpret_cast<const float2*>(sBe + n);
      const int ch = n >> 1;
      const float sc = sS[ch];
      float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t b0 = sW[(8 * nt + gid) * KP + 8 * ks + tig], b1 = sW[(8 * nt + gid) * KP + 8 * ks + tig + 4];
        if (X3) {
          const uint32_t l0 = sWl[(8 * nt + gid) * KP + 8 * ks + tig], l1 = sWl[(8 * nt + gid) * KP + 8 * ks + tig + 4];
          mma_tf32x3_16x8x8(c[0], ah[0][ks], al[0][X3 ? ks : 0], b0, b1, l0, l1);
          mma_tf32x3_16x8x8(c[1], ah[1][ks], al[X3 ? 1 : 0][X3 ? ks : 0], b0, b1, l0, l1);
        } else {
          mma_tf32_16x8x8(c[0], ah[0][ks], b0, b1);
          mma_tf32_16x8x8(c[1], ah[1][ks], b0, b1);
        }
      }
#pragma unroll
      for (int f = 0; f < 2; ++f) {
#pragma unroll
        for (int hi8 = 0; hi8 < 2; ++hi8) {
          const int q = 2 * f + hi8;
          const float val = fmaf((c[f][2 * hi8] - m2[q]) * r2[q], gam.x, bet.x);
          const float gate = fmaf((c[f][2 * hi8 + 1] - m2[q]) * r2[q], gam.y, bet.y);
          sT[(gid + 8 * q) * LDT + ch] = sc * val * bd_sigmoid(gate);
        }
      }
    }
    // the block's 48 channels of a row are one contiguous 192-byte run of x: float4 accesses, lane = consecutive float4
    __syncwarp();
    const long long left = M - m0;
    const int nrows = left < 32 ? (int)left : 32;
#pragma unroll
    for (int j0 = 0; j0 < CB / 4; j0 += 4) {       // 12 float4 per lane, 4 loads in flight
      float4 xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = lane + 32 * (j0 + u), row = i / (CB / 4), c4 = i - row * (CB / 4);
        if (row < nrows) xv[u] = *reinterpret_cast<const float4*>(x + (m0 + row) * C + c_lo + 4 * c4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = lane + 32 * (j0 + u), row = i / (CB / 4), c4 = i - row * (CB / 4);
        if (row < nrows) {
          const float4 d = *reinterpret_cast<const float4*>(sT + row * LDT + 4 * c4);
          *reinterpret_cast<float4*>(x + (m0 + row) * C + c_lo + 4 * c4) =
              make_float4(xv[u].x + d.x, xv[u].y + d.y, xv[u].z + d.z, xv[u].w + d.w);
        }
      }
    }
    __syncwarp();
  }
}

template <int HID, bool X3>
int launch_update_mma(const float* h, int ldh, const float* mr1, const float* g1, const float* be1, const float* w2t,
                      const float* b2, const float* mr2, const float* g2, const float* be2, const float* scale, float* x,
                      long long M, int C, long long rpi, int spi, cudaStream_t st) {
  constexpr int KS = (HID + 1 + 7) / 8, KP = 8 * KS + 4;
  const int nblk = C / 48;
  long long gx = ((M + 31) / 32 + 7) / 8;
  const long long cap = (148LL * 4 + nblk - 1) / nblk;          // ~4 CTAs per SM over all column blocks
  if (gx > cap) gx = cap;
  const int smem = ((X3 ? 2 : 1) * 96 * KP + 2 * 96 + 48 + 8 * 32 * 52) * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(dconv_update_mma_kernel<HID, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    bd_set_error("bd_dconv_expand_update: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return BD_ERR_CUDA;
  }
  dconv_update_mma_kernel<HID, X3><<<dim3((unsigned)gx, nblk), 256, smem, st>>>(h, ldh, mr1, g1, be1, w2t, b2, mr2, g2, be2,
                                                                           scale, x, M, C, rpi, spi);
  return bd_check_launch("dconv_update_mma_kernel");
}


// ---- the dilated k=3 convolution of a narrow DConv layer (hid 6) on mma.sync fragments -------------------------
// h[m, 0..5] = b1 + sum_tap W1[:, tap, :] . x[row of m shifted by (tap-1)*dil positions along the conv axis, :]
// and (sum, sumsq) of h per GroupNorm slab.  [rows x 3C] x [3C x 6]: the output is so narrow that a tcgen05 tile
// (N >= 16, its epilogue, a 16-wide padded h) costs more than the contraction.  Here the weights live in B
// fragments, a warp takes 32 rows, every lane reads float4 runs of the input row (the K index is permuted so that
// a lane's four fragment elements of two k-steps are one contiguous float4) and h is stored 8 floats wide.
template <int C, bool X3>
__global__ void __launch_bounds__(256, X3 ? 1 : 2) dconv_conv3_mma_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                              const float* __restrict__ b1, float* __restrict__ h,
                                                              double* __restrict__ sums, long long M, long long rpi,
                                                              int spi, int dil) {
  constexpr int HID = 6, KSC = C / 8, NP = C / 16;            // k-steps / float4 pairs per tap
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  // B fragment of k-step (tap, 2p + s): b0 <- channel 16p + 4tig + 2s, b1 <- channel 16p + 4tig + 2s + 1; column gid
  uint32_t bf[3 * KSC][2], bl[X3 ? 3 * KSC : 1][2];
#pragma unroll
  for (int tap = 0; tap < 3; ++tap)
#pragma unroll
    for (int ks = 0; ks < KSC; ++ks) {
      const int ch = 16 * (ks >> 1) + 4 * tig + 2 * (ks & 1);
      const float w0 = gid < HID ? __ldg(w1 + (size_t)gid * 3 * C + tap * C + ch) : 0.f;
      const float w1v = gid < HID ? __ldg(w1 + (size_t)gid * 3 * C + tap * C + ch + 1) : 0.f;
      if (X3) {
        split_tf32(w0, bf[tap * KSC + ks][0], bl[X3 ? tap * KSC + ks : 0][0]);
        split_tf32(w1v, bf[tap * KSC + ks][1], bl[X3 ? tap * KSC + ks : 0][1]);
      } else {
        bf[tap * KSC + ks][0] = to_tf32(w0);
        bf[tap * KSC + ks][1] = to_tf32(w1v);
      }
    }
  const float bz0 = 2 * tig < HID ? __ldg(b1 + 2 * tig) : 0.f, bz1 = 2 * tig + 1 < HID ? __ldg(b1 + 2 * tig + 1) : 0.f;
  const long long T = rpi / spi;
  const long long ntiles = (M + 31) / 32;
  const long long wstride = (long long)gridDim.x * 8;
  for (long long tile = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); tile < ntiles; tile += wstride) {
    const long long m0 = tile * 32;
    long long mrow[4];
    int tpos[4];
    bool live[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      mrow[q] = m0 + gid + 8 * q;
      live[q] = mrow[q] < M;
      tpos[q] = (int)(((live[q] ? mrow[q] : m0) % rpi) / spi);
    }
    float c[2][4] = {{bz0, bz1, bz0, bz1}, {bz0, bz1, bz0, bz1}};
#pragma unroll
    for (int tap = 0; tap < 3; ++tap) {
      float4 v[4][NP];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long long tt = tpos[q] + (long long)(tap - 1) * dil;
        const bool ok = live[q] && tt >= 0 && tt < T;
        const float* src = x + (mrow[q] + (long long)(tap - 1) * dil * spi) * C + 4 * tig;
#pragma unroll
        for (int p = 0; p < NP; ++p)
          v[q][p] = ok ? __ldg(reinterpret_cast<const float4*>(src + 16 * p)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int p = 0; p < NP; ++p) {
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          // rows 2f (gid + 16f) and 2f+1 (gid + 8 + 16f); a0/a1: k = tig of the two rows, a2/a3: k = tig + 4
          if (X3) {
            const float fe[4] = {v[2 * f][p].x, v[2 * f + 1][p].x, v[2 * f][p].y, v[2 * f + 1][p].y};
            const float fo[4] = {v[2 * f][p].z, v[2 * f + 1][p].z, v[2 * f][p].w, v[2 * f + 1][p].w};
            uint32_t eh[4], el[4], oh[4], ol[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              split_tf32(fe[e], eh[e], el[e]);
              split_tf32(fo[e], oh[e], ol[e]);
            }
            const int k0 = tap * KSC + 2 * p, kl0 = X3 ? k0 : 0, kl1 = X3 ? k0 + 1 : 0;
            mma_tf32x3_16x8x8(c[f], eh, el, bf[k0][0], bf[k0][1], bl[kl0][0], bl[kl0][1]);
            mma_tf32x3_16x8x8(c[f], oh, ol, bf[k0 + 1][0], bf[k0 + 1][1], bl[kl1][0], bl[kl1][1]);
          } else {
          const uint32_t a_even[4] = {to_tf32(v[2 * f][p].x), to_tf32(v[2 * f + 1][p].x), to_tf32(v[2 * f][p].y),
                                      to_tf32(v[2 * f + 1][p].y)};
          const uint32_t a_odd[4] = {to_tf32(v[2 * f][p].z), to_tf32(v[2 * f + 1][p].z), to_tf32(v[2 * f][p].w),
                                     to_tf32(v[2 * f + 1][p].w)};
          mma_tf32_16x8x8(c[f], a_even, bf[tap * KSC + 2 * p][0], bf[tap * KSC + 2 * p][1]);
          mma_tf32_16x8x8(c[f], a_odd, bf[tap * KSC + 2 * p + 1][0], bf[tap * KSC + 2 * p + 1][1]);
          }
        }
      }
    }
    // c[f] = {row gid+16f: cols 2tig, 2tig+1; row gid+8+16f: cols 2tig, 2tig+1}
    long long slab[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int f = q >> 1, hi8 = q & 1;
      const float h0 = c[f][2 * hi8], h1 = c[f][2 * hi8 + 1];
      if (live[q]) *reinterpret_cast<float2*>(h + mrow[q] * 8 + 2 * tig) = make_float2(h0, h1);
      float rs = h0 + h1, rq = fmaf(h0, h0, h1 * h1);          // padding columns are exactly zero
      rs += __shfl_xor_sync(0xffffffffu, rs, 1);
      rq += __shfl_xor_sync(0xffffffffu, rq, 1);
      rs += __shfl_xor_sync(0xffffffffu, rs, 2);
      rq += __shfl_xor_sync(0xffffffffu, rq, 2);
      const long long mm = live[q] ? mrow[q] : m0;
      slab[q] = (mm / rpi) * spi + (mm % spi);
      c[f][2 * hi8] = live[q] ? rs : 0.f;                        // reuse the accumulator registers for the row sums
      c[f][2 * hi8 + 1] = live[q] ? rq : 0.f;
    }
    // one slab for the whole tile (time branch, away from item boundaries): one atomic pair per warp
    const long long slab0 = __shfl_sync(0xffffffffu, slab[0], 0);
    const bool uniform = __all_sync(0xffffffffu, slab[0] == slab0 && slab[1] == slab0 && slab[2] == slab0 && slab[3] == slab0);
    if (uniform) {
      float ws = 0.f, wq = 0.f;
      if (tig == 0) {
        ws = (c[0][0] + c[0][2]) + (c[1][0] + c[1][2]);
        wq = (c[0][1] + c[0][3]) + (c[1][1] + c[1][3]);
      }
      ws = bd_warp_sum(ws);
      wq = bd_warp_sum(wq);
      if (lane == 0) {
        atomicAdd(&sums[2 * slab0], (double)ws);
        atomicAdd(&sums[2 * slab0 + 1], (double)wq);
      }
    } else if (tig == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (live[q]) {
          atomicAdd(&sums[2 * slab[q]], (double)c[q >> 1][2 * (q & 1)]);
          atomicAdd(&sums[2 * slab[q] + 1], (double)c[q >> 1][2 * (q & 1) + 1]);
        }
    }
  }
}


// ---- first encoder layer of each branch: Conv(k=8, s=4, p=2) over 2 / 4 input channels -> 48, GELU ---------------
// (hdemucs.py:110,139-144 with the input normalisation htdemucs.py:545-554 folded into the load).  K = 8 taps x C_in
// = 16 / 32 is too short for a tcgen05 tile; on mma.sync fragments the weights (48 x K) live in registers, a warp
// takes 32 output positions, and bias + GELU happen on the accumulator fragment.
//   CM = false (frequency branch): x [B, I1, Jin, 4] channels-last, the window of output i0 is the 32 contiguous
//        floats starting at position 4*i0 - 2; K index permuted so that a lane reads float4 = one position.
//   CM = true  (time branch):      x [B, 2, Jin] channel-major (the raw mix), k-step = channel, k = tap.
// Positions outside [0, Jin) read as zero AFTER the normalisation (zero padding of the normalised tensor).
template <int CIN, bool CM, bool X3>
__global__ void __launch_bounds__(256, X3 ? 1 : 2) conv_first_mma_kernel(const float* __restrict__ x, const float* __restrict__ norm,
                                                                int norm_stride, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ out,
                                                                long long M, int I1, int Io, int Jin) {
  constexpr int K = 8 * CIN, KS = K / 8, NT = 6, COUT = 48;
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  // logical k of fragment element (ks, tig, half) -> (tap, channel)
  //   CM:  ks = channel, tap = tig + 4*half
  //   !CM: pair p = ks>>1, position-in-window = 4p + tig (one float4 = 4 channels), channel = 2*(ks&1) + half
  uint32_t bf[KS][NT][2], bl[X3 ? KS : 1][NT][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int tap = CM ? tig + 4 * half : 4 * (ks >> 1) + tig;
        const int ch = CM ? ks : 2 * (ks & 1) + half;
        const float wv = __ldg(w + (size_t)(8 * nt + gid) * K + tap * CIN + ch);
        if (X3) split_tf32(wv, bf[ks][nt][half], bl[X3 ? ks : 0][nt][half]);
        else bf[ks][nt][half] = to_tf32(wv);
      }
  float bz[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    bz[nt][0] = __ldg(bias + 8 * nt + 2 * tig);
    bz[nt][1] = __ldg(bias + 8 * nt + 2 * tig + 1);
  }
  const long long rows_per_item = (long long)I1 * Io;
  const long long ntiles = (M + 31) / 32;
  const long long wstride = (long long)gridDim.x * 8;
  for (long long tile = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); tile < ntiles; tile += wstride) {
    const long long m0 = tile * 32;
    uint32_t a[2][KS][4], al[X3 ? 2 : 1][X3 ? KS : 1][4];
    auto put = [&](int f, int ks, int e, float v) {
      if (X3) split_tf32(v, a[f][ks][e], al[X3 ? f : 0][X3 ? ks : 0][e]);
      else a[f][ks][e] = to_tf32(v);
    };
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long m = m0 + gid + 8 * q;
      const long long mm = m < M ? m : m0;
      const int b = (int)(mm / rows_per_item);
      const long long r = mm - (long long)b * rows_per_item;
      const int i1 = (int)(r / Io), i0 = (int)(r - (long long)i1 * Io);
      const float mean = __ldg(norm + (size_t)b * norm_stride), rstd = __ldg(norm + (size_t)b * norm_stride + 2);
      const int f = q >> 1, hi8 = q & 1;
      if (CM) {
        const float* src = x + (size_t)b * CIN * Jin;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int pos = 4 * i0 - 2 + tig + 4 * half;
            const float v = (pos >= 0 && pos < Jin) ? (__ldg(src + (size_t)ks * Jin + pos) - mean) * rstd : 0.f;
            put(f, ks, 2 * half + hi8, v);
          }
      } else {
        const float* src = x + (((size_t)b * I1 + i1) * Jin) * CIN;
#pragma unroll
        for (int p = 0; p < KS / 2; ++p) {
          const int pos = 4 * i0 - 2 + 4 * p + tig;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (pos >= 0 && pos < Jin) {
            v = __ldg(reinterpret_cast<const float4*>(src + (size_t)pos * CIN));
            v = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
          }
          put(f, 2 * p, hi8, v.x);                  // ks = 2p:   channels 0 (half 0), 1 (half 1)
          put(f, 2 * p, 2 + hi8, v.y);
          put(f, 2 * p + 1, hi8, v.z);              // ks = 2p+1: channels 2, 3
          put(f, 2 * p + 1, 2 + hi8, v.w);
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        float c[4] = {bz[nt][0], bz[nt][1], bz[nt][0], bz[nt][1]};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          if (X3)
            mma_tf32x3_16x8x8(c, a[f][ks], al[X3 ? f : 0][X3 ? ks : 0], bf[ks][nt][0], bf[ks][nt][1], bl[X3 ? ks : 0][nt][0],
                              bl[X3 ? ks : 0][nt][1]);
          else
            mma_tf32_16x8x8(c, a[f][ks], bf[ks][nt][0], bf[ks][nt][1]);
        }
#pragma unroll
        for (int hi8 = 0; hi8 < 2; ++hi8) {
          const long long m = m0 + gid + 8 * (2 * f + hi8);
          if (m < M)
            *reinterpret_cast<float2*>(out + m * COUT + 8 * nt + 2 * tig) =
                make_float2(bd_gelu(c[2 * hi8]), bd_gelu(c[2 * hi8 + 1]));
        }
      }
    }
  }
}

}  // namespace

extern "C" {

int bd_encoder_conv0(const float* x, int channel_major, const float* norm, int norm_stride, const float* w, const float* bias,
                     float* out, int B, int I1, int Io, int Jin, int cin, int cout, int math, void* stream) {
  BD_REQUIRE(cout == 48 && ((channel_major && cin == 2 && I1 == 1) || (!channel_major && cin == 4)),
             "bd_encoder_conv0: only the htdemucs first layers are built (cin=%d cout=%d channel_major=%d)", cin, cout, channel_major);
  BD_REQUIRE(B > 0 && I1 > 0 && Io > 0 && Jin > 0 && 4 * (Io - 1) - 2 < Jin, "bd_encoder_conv0: bad sizes");
  BD_REQUIRE((((uintptr_t)x | (uintptr_t)out) & 15) == 0, "bd_encoder_conv0: unaligned tensor");
  const long long M = (long long)B * I1 * Io;
  long long grid = ((M + 31) / 32 + 7) / 8;
  if (grid > 148LL * 2 * 4) grid = 148LL * 2 * 4;
  const bool x3 = math == BD_MATH_TF32X3 || math == BD_MATH_BF16X3;
  const cudaStream_t st = (cudaStream_t)stream;
  if (channel_major) {
    if (x3) conv_first_mma_kernel<2, true, true><<<(unsigned)grid, 256, 0, st>>>(x, norm, norm_stride, w, bias, out, M, I1, Io, Jin);
    else conv_first_mma_kernel<2, true, false><<<(unsigned)grid, 256, 0, st>>>(x, norm, norm_stride, w, bias, out, M, I1, Io, Jin);
  } else {
    if (x3) conv_first_mma_kernel<4, false, true><<<(unsigned)grid, 256, 0, st>>>(x, norm, norm_stride, w, bias, out, M, I1, Io, Jin);
    else conv_first_mma_kernel<4, false, false><<<(unsigned)grid, 256, 0, st>>>(x, norm, norm_stride, w, bias, out, M, I1, Io, Jin);
  }
  return bd_check_launch("conv_first_mma_kernel");
}

int bd_dconv_conv3(const float* x, const float* w1, const float* b1, float* h, int ldh, double* sums1, long long M, int C,
                   int hid, long long rows_per_item, int slabs_per_item, int dilation, int math, void* stream) {
  BD_REQUIRE(hid == 6 && C == 48 && ldh == 8, "bd_dconv_conv3: only hid 6 / C 48 / ldh 8 is built (hid=%d C=%d ldh=%d)", hid, C, ldh);
  BD_REQUIRE(M > 0 && rows_per_item > 0 && slabs_per_item > 0 && rows_per_item % slabs_per_item == 0 && dilation > 0,
             "bd_dconv_conv3: bad sizes");
  BD_REQUIRE((((uintptr_t)x | (uintptr_t)h) & 15) == 0, "bd_dconv_conv3: unaligned tensor");
  long long grid = ((M + 31) / 32 + 7) / 8;
  if (grid > 148LL * 2 * 4) grid = 148LL * 2 * 4;     // 2 resident CTAs per SM, 4 rounds
  if (math == BD_MATH_TF32X3 || math == BD_MATH_BF16X3)
    dconv_conv3_mma_kernel<48, true><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, w1, b1, h, sums1, M, rows_per_item,
                                                                                       slabs_per_item, dilation);
  else
    dconv_conv3_mma_kernel<48, false><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, w1, b1, h, sums1, M, rows_per_item,
                                                                                        slabs_per_item, dilation);
  return bd_check_launch("dconv_conv3_mma_kernel");
}

int bd_dconv_expand_stats(const float* h, int ldh, int hid, const float* mean_rstd1, const float* gamma1,
                          const float* beta1, const float* w2t, const float* b2, double* sums2, double* gram_ws,
                          long long M, int C, long long rows_per_item, int slabs_per_item, void* stream) {
  BD_REQUIRE(hid > 0 && hid <= MAX_HID && C % 2 == 0 && ldh >= hid && M > 0, "bd_dconv_expand_stats: bad sizes (hid=%d C=%d)", hid, C);
  if (gram_ws && M % rows_per_item == 0 && gram_supported(hid, ldh, rows_per_item, slabs_per_item)) {
    return launch_gram(h, ldh, hid, mean_rstd1, gamma1, beta1, w2t, b2, sums2, gram_ws, M, C, rows_per_item,
                       slabs_per_item, (cudaStream_t)stream);
  }
  return launch_expand<false>(h, ldh, hid, mean_rstd1, gamma1, beta1, w2t, b2, sums2, nullptr, nullptr, nullptr, nullptr,
                              nullptr, M, C, rows_per_item, slabs_per_item, (cudaStream_t)stream);
}

int bd_dconv_expand_update(const float* h, int ldh, int hid, const float* mean_rstd1, const float* gamma1,
                           const float* beta1, const float* w2t, const float* b2, const float* mean_rstd2,
                           const float* gamma2, const float* beta2, const float* scale, float* x, long long M, int C,
                           long long rows_per_item, int slabs_per_item, int math, void* stream) {
  BD_REQUIRE(hid > 0 && hid <= MAX_HID && C % 2 == 0 && ldh >= hid && M > 0, "bd_dconv_expand_update: bad sizes (hid=%d C=%d)", hid, C);
#define BD_MMA_ARGS h, ldh, mean_rstd1, gamma1, beta1, w2t, b2, mean_rstd2, gamma2, beta2, scale, x, M, C, rows_per_item, slabs_per_item, (cudaStream_t)stream
  if ((math == BD_MATH_TF32 || math == BD_MATH_BF16) && C % 48 == 0 && (hid == 6 || hid == 12 || hid == 24 || hid == 48)) {
    if (hid == 6) return launch_update_mma<6, false>(BD_MMA_ARGS);
    if (hid == 12) return launch_update_mma<12, false>(BD_MMA_ARGS);
    if (hid == 24) return launch_update_mma<24, false>(BD_MMA_ARGS);
    return launch_update_mma<48, false>(BD_MMA_ARGS);
  }
  // strict modes: the same kernel with hi/lo operand splits in registers (three mma.sync per product) where it beats
  // the exact FFMA kernel below (measured: hid 6 1.31 vs 1.74 ms; hid 12 1.30 vs 1.08 ms per 4 launches at 16 segments)
  if ((math == BD_MATH_TF32X3 || math == BD_MATH_BF16X3) && C % 48 == 0 && hid == 6)
    return launch_update_mma<6, true>(BD_MMA_ARGS);
#undef BD_MMA_ARGS
  return launch_expand<true>(h, ldh, hid, mean_rstd1, gamma1, beta1, w2t, b2, nullptr, mean_rstd2, gamma2, beta2, scale, x,
                             M, C, rows_per_item, slabs_per_item, (cudaStream_t)stream);
}

}  // extern "C"

// K1 / K2: fused STFT and iSTFT for the complex-as-channels spectrogram path.
//
// Replaces reference spec.py:11-47 (spectro / ispectro) together with
// HTDemucs._spec/_magnitude (htdemucs.py:420-461) and _mask/_ispec (:442-471) and the
// surrounding pad1d reflect padding (hdemucs.py:23-40), per SURVEY.md section 8a rows 1-5,15-17.
//
// One CTA = one frame.  The two audio channels of a frame are packed as the real and
// imaginary part of ONE 4096-point complex FFT (Stockham auto-sort, radix 8, 4 passes
// through shared memory) and separated with the Hermitian-symmetry identity.  HBM-bound by
// design: the signal is read once (the 4x frame overlap is served by L2) and the packed
// spectrogram is written once, position-innermost ([B, T, F, 4] floats = one float4 per bin).
#include "common.cuh"

namespace {

constexpr int NFFT = 4096;
constexpr int NTHR = 512;   // one radix-8 butterfly per thread per pass
constexpr int SPAD = NFFT + NFFT / 8;

__device__ __forceinline__ int sidx(int i) { return i + (i >> 3); }  // bank-conflict padding

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by SIGN * i
template <int SIGN>
__device__ __forceinline__ float2 cmul_i(float2 a) {
  return SIGN > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// radix-4 DIF, outputs in natural order
template <int SIGN>
__device__ __forceinline__ void fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 c0 = cadd(a0, a2), c1 = cadd(a1, a3), d0 = csub(a0, a2), d1 = cmul_i<SIGN>(csub(a1, a3));
  a0 = cadd(c0, c1);
  a2 = csub(c0, c1);
  a1 = cadd(d0, d1);
  a3 = csub(d0, d1);
}

// radix-8 DIF DFT with kernel exp(SIGN * 2 pi i n k / 8); result left in natural order in v[]
template <int SIGN>
__device__ __forceinline__ void fft8(float2* v) {
  const float h = 0.70710678118654752440f;
  float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
  float2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
  b1 = cmul(b1, make_float2(h, SIGN * h));
  b2 = cmul_i<SIGN>(b2);
  b3 = cmul(b3, make_float2(-h, SIGN * h));
  fft4<SIGN>(a0, a1, a2, a3);
  fft4<SIGN>(b0, b1, b2, b3);
  v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
  v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// 4096-point complex FFT of the values held as v[r] = x[j + 512 r] by thread j.
// tw[m] = (cos, -sin)(2 pi m / 4096); SIGN=+1 conjugates it.  Result: natural order in (sre, sim).
template <int SIGN>
__device__ __forceinline__ void fft4096(float2* v, float* sre, float* sim, const float2* __restrict__ tw) {
  const int j = threadIdx.x;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int Ns = 1 << (3 * pass);
    const int k = j & (Ns - 1);
    if (pass > 0) {
      const int step = k * (NTHR / Ns);
#pragma unroll
      for (int r = 1; r < 8; ++r) {
        float2 w = __ldg(&tw[step * r]);
        if (SIGN > 0) w.y = -w.y;
        v[r] = cmul(v[r], w);
      }
    }
    fft8<SIGN>(v);
    const int j0 = ((j - k) << 3) + k;
    if (pass > 0) __syncthreads();  // everyone has finished reading the previous pass
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int p = sidx(j0 + r * Ns);
      sre[p] = v[r].x;
      sim[p] = v[r].y;
    }
    __syncthreads();
    if (pass < 3) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        int p = sidx(j + r * NTHR);
        v[r] = make_float2(sre[p], sim[p]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K1. mix [B, 2, L] planar -> spec [B, T, 2048, 4] with channel order (c0.re, c0.im, c1.re, c1.im).
// Also accumulates, per item, (sum, sumsq) of the spectrogram values and of the raw samples:
// the statistics of the per-item normalisation at htdemucs.py:545-554.
__global__ void __launch_bounds__(NTHR) stft_cac_kernel(const float* __restrict__ mix, const float* __restrict__ win,
                                                        const float2* __restrict__ tw, float* __restrict__ spec,
                                                        double* __restrict__ stats, int L, int T) {
  __shared__ float sre[SPAD];
  __shared__ float sim[SPAD];
  __shared__ float red[4][16];
  const int t = blockIdx.x, b = blockIdx.y, j = threadIdx.x;
  const float* x0 = mix + (size_t)b * 2 * L;
  const float* x1 = x0 + L;
  float2 v[8];
  float ts = 0.f, tq = 0.f;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int n = j + r * NTHR;
    int src = t * 1024 + n - 1536;           // hop * t - 3 hop / 2 (htdemucs.py:433-434)
    const bool own = (r == 3 || r == 4) && src < L;  // the hop block this frame owns
    src = src < 0 ? -src : src;
    src = src >= L ? 2 * (L - 1) - src : src;
    const float a = __ldg(x0 + src), c = __ldg(x1 + src), w = __ldg(win + n);
    if (own) {
      ts += a + c;
      tq = fmaf(a, a, fmaf(c, c, tq));
    }
    v[r] = make_float2(a * w, c * w);
  }
  fft4096<-1>(v, sre, sim, tw);
  float4* out = reinterpret_cast<float4*>(spec) + ((size_t)b * T + t) * 2048;
  float fs = 0.f, fq = 0.f;
  const float sc = 1.0f / 128.0f;  // 1/2 (channel split) * 1/sqrt(4096) (normalized=True)
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int f = j + r * NTHR;
    const int pf = sidx(f), pm = sidx((NFFT - f) & (NFFT - 1));
    const float zr = sre[pf], zi = sim[pf], mr = sre[pm], mi = sim[pm];
    float4 o = make_float4((zr + mr) * sc, (zi - mi) * sc, (zi + mi) * sc, (mr - zr) * sc);
    out[f] = o;
    fs += (o.x + o.y) + (o.z + o.w);
    fq = fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, fmaf(o.w, o.w, fq))));
  }
  // one frame's partial sums (<= 8192 values) in fp32, frames combined in fp64
  fs = bd_warp_sum(fs); fq = bd_warp_sum(fq); ts = bd_warp_sum(ts); tq = bd_warp_sum(tq);
  if ((j & 31) == 0) {
    red[0][j >> 5] = fs; red[1][j >> 5] = fq; red[2][j >> 5] = ts; red[3][j >> 5] = tq;
  }
  __syncthreads();
  if (j < 4) {
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < 16; ++w) acc += (double)red[j][w];
    atomicAdd(&stats[b * 4 + j], acc);
  }
}

// (sum, sumsq) -> (mean, std, 1/(1e-5+std), 0) with the UNBIASED std of Tensor.std() (htdemucs.py:546,553)
__global__ void finalize_item_norm_kernel(const double* __restrict__ stats, float* __restrict__ norm, int B,
                                          double n_freq, double n_time) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * B) return;
  int b = i >> 1, which = i & 1;
  double n = which ? n_time : n_freq;
  double s = stats[b * 4 + 2 * which], q = stats[b * 4 + 2 * which + 1];
  double mean = s / n;
  double var = (q - s * mean) / (n - 1.0);
  float sd = (float)sqrt(var > 0.0 ? var : 0.0);
  float* o = norm + (size_t)b * 8 + which * 4;
  o[0] = (float)mean;
  o[1] = sd;
  o[2] = 1.0f / (1e-5f + sd);
  o[3] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// K2a. spec [B, T, 2048, 4S] (channel = 4s + 2c + {re,im}) -> windowed time frames [B, S, 2, T, 4096].
// De-normalisation x*std+mean (htdemucs.py:624-626), CaC unpack (:463-471), zero Nyquist and edge frames
// (:444-445), inverse FFT, synthesis window and the 1/1.5 envelope are fused.
__global__ void __launch_bounds__(NTHR) istft_frames_kernel(const float* __restrict__ spec, const float* __restrict__ norm,
                                                            const float* __restrict__ win, const float2* __restrict__ tw,
                                                            float* __restrict__ frames, int S, int T) {
  __shared__ float sre[SPAD];
  __shared__ float sim[SPAD];
  const int t = blockIdx.x, b = blockIdx.y, j = threadIdx.x;
  const float mean = norm[b * 8 + 0], sd = norm[b * 8 + 1];
  const float4* in = reinterpret_cast<const float4*>(spec) + ((size_t)b * T + t) * 2048 * S;
  for (int s = 0; s < S; ++s) {
    float2 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int f = j + r * NTHR;
      if (f == 2048) {
        v[r] = make_float2(0.f, 0.f);
        continue;
      }
      const bool mirror = f > 2048;
      const int g = mirror ? NFFT - f : f;
      float4 q = __ldg(in + (size_t)g * S + s);
      q.x = fmaf(q.x, sd, mean); q.y = fmaf(q.y, sd, mean);
      q.z = fmaf(q.z, sd, mean); q.w = fmaf(q.w, sd, mean);
      if (g == 0) { q.y = 0.f; q.w = 0.f; }   // c2r transform ignores imag(DC)
      // Z = X0 + i X1 on the lower half, conj(X0) + i conj(X1) on the mirrored half
      v[r] = mirror ? make_float2(q.x + q.w, q.z - q.y) : make_float2(q.x - q.w, q.y + q.z);
    }
    fft4096<1>(v, sre, sim, tw);
    float* f0 = frames + ((((size_t)b * S + s) * 2 + 0) * T + t) * NFFT;
    float* f1 = f0 + (size_t)T * NFFT;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int n = j + r * NTHR;
      const float w = __ldg(win + n) * (1.0f / 64.0f / 1.5f);
      const int p = sidx(n);
      f0[n] = sre[p] * w;
      f1[n] = sim[p] * w;
    }
    __syncthreads();
  }
}

// K2b. overlap-add of the 4 frames covering each sample, crop (htdemucs.py:449), and the fused tail
// out = xt*stdt + meant + x (htdemucs.py:653-657).  xt is the time decoder output, channels-last
// [B, Lseg, 2S] with channel = 2s + c.
__global__ void ola_combine_kernel(const float* __restrict__ frames, const float* __restrict__ xt,
                                   const float* __restrict__ norm, float* __restrict__ out, int S, int T, int Lseg,
                                   int Lout) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (n >= Lout) return;
  const float meant = norm[b * 8 + 4], stdt = norm[b * 8 + 5];
  const int p = n + 1536;
  const int t_hi = min(p >> 10, T - 1);
  const int t_lo = max(0, (p - (NFFT - 1) + 1023) >> 10);
  const float* xrow = xt ? xt + ((size_t)b * Lseg + n) * 2 * S : nullptr;
  for (int sc = 0; sc < 2 * S; ++sc) {
    const float* fr = frames + ((size_t)b * 2 * S + sc) * T * NFFT;
    float acc = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) acc += fr[(size_t)t * NFFT + (p - (t << 10))];
    if (xrow) acc += fmaf(xrow[sc], stdt, meant);
    out[((size_t)b * 2 * S + sc) * Lout + n] = acc;
  }
}

// K2 fused: spec [B, T, S, 2048, 4] (source-major, written that way by the last decoder layer) ->
// out [B, S, 2, Lout].  One CTA owns NH consecutive hop blocks of one (item, source): it walks the NH + 3
// frames that touch them, inverse-transforms each (two channels per complex FFT), and overlap-adds into a
// 4096-sample ring in shared memory; a hop block leaves the SM exactly once, already cropped, with the
// time-branch output and the de-normalisation folded in.  No frames buffer in HBM: the algorithmic traffic
// (spectrogram in, waveform out) is the only traffic, bar the 3 recomputed halo frames per chunk.
constexpr int NH = 24;

__global__ void __launch_bounds__(NTHR, 2) istft_ola_kernel(const float* __restrict__ spec, const float* __restrict__ norm,
                                                         const float* __restrict__ win, const float2* __restrict__ tw,
                                                         const float* __restrict__ xt, float* __restrict__ out, int S,
                                                         int T, int Lseg, int Lout) {
  extern __shared__ __align__(16) float dsm[];
  float* sre = dsm;
  float* sim = sre + SPAD;
  float* ring0 = sim + SPAD;
  float* ring1 = ring0 + NFFT;
  const int h0 = blockIdx.x * NH, s = blockIdx.y, b = blockIdx.z, j = threadIdx.x;
  const float mean = norm[b * 8 + 0], sd = norm[b * 8 + 1];
  const float meant = norm[b * 8 + 4], stdt = norm[b * 8 + 5];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    ring0[j + r * NTHR] = 0.f;
    ring1[j + r * NTHR] = 0.f;
  }
  const int h_end = min(h0 + NH, T + 2);           // hop blocks [h0, h_end)
  for (int t = h0 - 3; t < h_end; ++t) {
    if (t >= 0 && t < T) {
      const float4* in = reinterpret_cast<const float4*>(spec) + (((size_t)b * T + t) * S + s) * 2048;
      float2 v[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int f = j + r * NTHR;
        if (f == 2048) {
          v[r] = make_float2(0.f, 0.f);
          continue;
        }
        const bool mirror = f > 2048;
        const int g = mirror ? NFFT - f : f;
        float4 q = __ldg(in + g);
        q.x = fmaf(q.x, sd, mean); q.y = fmaf(q.y, sd, mean);
        q.z = fmaf(q.z, sd, mean); q.w = fmaf(q.w, sd, mean);
        if (g == 0) { q.y = 0.f; q.w = 0.f; }
        v[r] = mirror ? make_float2(q.x + q.w, q.z - q.y) : make_float2(q.x - q.w, q.y + q.z);
      }
      fft4096<1>(v, sre, sim, tw);                 // ends with a __syncthreads
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int n = j + r * NTHR;
        const float w = __ldg(win + n) * (1.0f / 64.0f / 1.5f);
        const int p = sidx(n);
        const int pos = t * 1024 + n;                 // position on the overlap-add grid
        const int ri = pos & (NFFT - 1);              // each thread owns its ring slots: no races
        if (pos >= h0 * 1024) {                       // halo frames only feed the hop blocks this CTA owns
          ring0[ri] += sre[p] * w;
          ring1[ri] += sim[p] * w;
        }
      }
    }
    __syncthreads();
    if (t >= h0) {                                 // hop block t has now received frames t-3 .. t
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = j + r * NTHR;
        const int ri = (t * 1024 + i) & (NFFT - 1);
        const int n = t * 1024 + i - 1536;
        const float a0 = ring0[ri], a1 = ring1[ri];
        ring0[ri] = 0.f;
        ring1[ri] = 0.f;
        if (n >= 0 && n < Lout) {
          float t0 = 0.f, t1 = 0.f;
          if (xt) {
            const float2 x2 = __ldg(reinterpret_cast<const float2*>(xt + ((size_t)b * Lseg + n) * 2 * S + 2 * s));
            t0 = fmaf(x2.x, stdt, meant);
            t1 = fmaf(x2.y, stdt, meant);
          }
          out[(((size_t)b * S + s) * 2 + 0) * Lout + n] = a0 + t0;
          out[(((size_t)b * S + s) * 2 + 1) * Lout + n] = a1 + t1;
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" {

int bd_stft_cac(const float* mix, const float* window, const float* twiddle, float* spec, double* stats,
                int B, int A, int L, void* stream) {
  BD_REQUIRE(A == 2, "bd_stft_cac: audio_channels must be 2 (got %d)", A);
  BD_REQUIRE(B > 0 && L >= 4096, "bd_stft_cac: need B > 0 and L >= 4096 (got B=%d L=%d)", B, L);
  int T = (L + 1023) / 1024;
  BD_REQUIRE(L > 1536 + (T * 1024 - L), "bd_stft_cac: reflect padding longer than the signal");
  stft_cac_kernel<<<dim3(T, B), NTHR, 0, (cudaStream_t)stream>>>(mix, window, (const float2*)twiddle, spec, stats, L, T);
  return bd_check_launch("stft_cac_kernel");
}

int bd_finalize_item_norm(const double* stats, float* norm, int B, double n_freq, double n_time, void* stream) {
  finalize_item_norm_kernel<<<bd_cdiv(2 * B, 64), 64, 0, (cudaStream_t)stream>>>(stats, norm, B, n_freq, n_time);
  return bd_check_launch("finalize_item_norm_kernel");
}

int bd_istft_frames(const float* spec, const float* norm, const float* window, const float* twiddle, float* frames,
                    int B, int S, int T, void* stream) {
  BD_REQUIRE(B > 0 && S > 0 && T > 0, "bd_istft_frames: bad sizes");
  istft_frames_kernel<<<dim3(T, B), NTHR, 0, (cudaStream_t)stream>>>(spec, norm, window, (const float2*)twiddle, frames,
                                                                    S, T);
  return bd_check_launch("istft_frames_kernel");
}

int bd_istft_ola(const float* spec, const float* norm, const float* window, const float* twiddle, const float* xt,
                 float* out, int B, int S, int T, int Lseg, int Lout, void* stream) {
  BD_REQUIRE(B > 0 && S > 0 && T > 0 && Lout > 0 && Lout <= Lseg, "bd_istft_ola: bad sizes");
  constexpr int smem = (2 * SPAD + 2 * NFFT) * (int)sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(istft_ola_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    bd_set_error("bd_istft_ola: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return BD_ERR_CUDA;
  }
  istft_ola_kernel<<<dim3((T + 2 + NH - 1) / NH, S, B), NTHR, smem, (cudaStream_t)stream>>>(
      spec, norm, window, (const float2*)twiddle, xt, out, S, T, Lseg, Lout);
  return bd_check_launch("istft_ola_kernel");
}

int bd_ola_combine(const float* frames, const float* xt, const float* norm, float* out, int B, int S, int T, int Lseg,
                   int Lout, void* stream) {
  BD_REQUIRE(Lout > 0 && Lout <= Lseg, "bd_ola_combine: Lout must be in (0, Lseg]");
  ola_combine_kernel<<<dim3(bd_cdiv(Lout, 256), B), 256, 0, (cudaStream_t)stream>>>(frames, xt, norm, out, S, T, Lseg,
                                                                                  Lout);
  return bd_check_launch("ola_combine_kernel");
}

}  // extern "C"

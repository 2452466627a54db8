// Kernels that only Hybrid Demucs v3 needs (reference demucs/hdemucs.py:123-157,304-335; demucs/demucs.py:20-67,157-216):
//   * GroupNorm with G > 1 groups over a channels-last tensor [B, rows, C] (norm_groups = 4 in the two innermost layers):
//     statistics per (item, group), then affine + GELU / GLU (natural channel order: value c, gate c + C/2) with an
//     optional row crop, because HDecLayer normalises the UNCROPPED transposed-convolution output (hdemucs.py:326-331);
//   * the BLSTM of the DConv branch: frame splitting (max_steps = 200, stride 100), the recurrence of one
//     bidirectional layer from pre-computed input projections, and the un-framing with the skip connection;
//   * LocalState: 4-head attention with learned per-query decay and no self reference.
// These layers sit on 1/16 .. 1/32 of the time resolution (T <= 336 rows per item): they are latency-, not
// bandwidth-bound, and are written for clarity in fp32; the convolutions around them go through bd_conv_gemm.
#include <stdlib.h>
#include "common.cuh"
#include "../../include/demucs_b200.h"

namespace {

// ---- GroupNorm(G) ------------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums, long long rows, int C, int G) {
  __shared__ double red[64];
  const int b = blockIdx.y / G, g = blockIdx.y % G, cg = C / G;
  const float* xb = x + (size_t)b * rows * C + (size_t)g * cg;
  const long long n = rows * cg;
  double s = 0.0, q = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cg;
    const float v = __ldg(xb + r * C + (i - r * cg));
    s += v;
    q += (double)v * v;
  }
  bd_block_sum2(s, q, red);
  if (threadIdx.x == 0) {
    atomicAdd(&sums[2 * blockIdx.y], s);
    atomicAdd(&sums[2 * blockIdx.y + 1], q);
  }
}

// y[b, r, c] = act(gn(x[b, row0 + r, :]))[c]; act: NONE / GELU keep C channels, GLU gives C/2 (a * sigmoid(gate))
__global__ void gn_act_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ mr,
                              const float* __restrict__ gamma, const float* __restrict__ beta,
                              const float* __restrict__ addend, long long rows_in, long long row0, long long rows_out,
                              int C, int G, int act, long long y_item_stride) {
  const int Co = act == BD_ACT_GLU ? C / 2 : C;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= rows_out * Co) return;
  const long long r = i / Co;
  const int c = (int)(i - r * Co), cg = C / G;
  const float* xr = x + ((size_t)b * rows_in + row0 + r) * C;
  auto norm = [&](int ch) {
    const int g = ch / cg;
    const float mean = __ldg(mr + 2 * (b * G + g)), rstd = __ldg(mr + 2 * (b * G + g) + 1);
    return fmaf((__ldg(xr + ch) - mean) * rstd, __ldg(gamma + ch), __ldg(beta + ch));
  };
  float v = norm(c);
  if (act == BD_ACT_GELU) v = bd_gelu(v);
  else if (act == BD_ACT_GLU) v = v * (1.0f / (1.0f + expf(-norm(c + Co))));
  const size_t at = (size_t)b * y_item_stride + r * Co + c;
  y[at] = addend ? v + __ldg(addend + at) : v;      // the next decoder layer's skip (x + skip, hdemucs.py:310)
}

// ---- BLSTM ---------------------------------------------------------------------------------------------------------
// frames[(b*nf + k), j, c] = x[b, k*stride + j, c] (zero past the end of the item): utils.unfold (utils.py:20-35)
__global__ void frame_kernel(const float* __restrict__ x, float* __restrict__ frames, long long T, int C, int nf, int width,
                             int stride) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int bk = blockIdx.y, b = bk / nf, k = bk % nf;
  if (i >= (long long)width * C) return;
  const long long j = i / C;
  const long long t = (long long)k * stride + j;
  frames[(size_t)bk * width * C + i] = t < T ? __ldg(x + ((size_t)b * T + t) * C + (i - j * C)) : 0.f;
}

// out[b, t, c] = frames[b*nf + k(t), t - k(t)*stride, c] + skip[b, t, c]: frame 0 keeps its first width - limit
// samples, middle frames their central `stride`, the last frame everything after `limit` (demucs.py:52-64)
__global__ void unframe_add_kernel(const float* __restrict__ frames, const float* __restrict__ skip, float* __restrict__ out,
                                   long long T, int C, int nf, int width, int stride) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= T * C) return;
  const long long t = i / C;
  const int limit = stride / 2;
  long long k = t < limit ? 0 : (t - limit) / stride;
  if (k > nf - 1) k = nf - 1;
  const float v = __ldg(frames + (((size_t)b * nf + k) * width + (t - k * stride)) * C + (i - t * C));
  out[(size_t)b * T * C + i] = v + __ldg(skip + (size_t)b * T * C + i);
}

// One time step of both directions of one LSTM layer.  pre [N, T, 2, 4H] holds x W_ih^T + b_ih + b_hh (direction-
// major, gate order i f g o as nn.LSTM), whhT [2, H, 4H] the transposed recurrent weights (coalesced over units),
// hbuf [2 (ping-pong), 2, N, H], cbuf [2, N, H], out [N, T, 2H] = [forward | backward].
// Block = 64 hidden units x 4 batch items of one direction; the previous hidden vectors sit in shared memory.
__global__ void __launch_bounds__(256) lstm_step_kernel(const float* __restrict__ pre, const float* __restrict__ whhT,
                                                        float* __restrict__ hbuf, float* __restrict__ cbuf,
                                                        float* __restrict__ out, int N, int T, int H, int step) {
  extern __shared__ float sh[];                       // [4][H]
  const int dir = blockIdx.z, t = dir ? T - 1 - step : step;
  const int u = blockIdx.x * 64 + (threadIdx.x & 63), bl = threadIdx.x >> 6, b = blockIdx.y * 4 + bl;
  const float* hprev = hbuf + ((size_t)((step & 1) * 2 + dir) * N) * H;
  float* hnext = hbuf + ((size_t)(((step + 1) & 1) * 2 + dir) * N) * H;
  for (int i = threadIdx.x; i < 4 * H; i += 256) {
    const int bb = blockIdx.y * 4 + i / H;
    sh[i] = bb < N ? hprev[(size_t)bb * H + (i % H)] : 0.f;
  }
  __syncthreads();
  if (u >= H || b >= N) return;
  const float* w = whhT + (size_t)dir * H * 4 * H + u;
  const float* hp = sh + bl * H;
  float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f;
#pragma unroll 4
  for (int k = 0; k < H; ++k) {
    const float hk = hp[k];
    const float* wk = w + (size_t)k * 4 * H;
    gi = fmaf(hk, __ldg(wk), gi);
    gf = fmaf(hk, __ldg(wk + H), gf);
    gg = fmaf(hk, __ldg(wk + 2 * H), gg);
    go = fmaf(hk, __ldg(wk + 3 * H), go);
  }
  const float* p = pre + (((size_t)b * T + t) * 2 + dir) * 4 * H + u;
  gi += __ldg(p); gf += __ldg(p + H); gg += __ldg(p + 2 * H); go += __ldg(p + 3 * H);
  const size_t ci = ((size_t)dir * N + b) * H + u;
  const float si = 1.f / (1.f + expf(-gi)), sf = 1.f / (1.f + expf(-gf)), so = 1.f / (1.f + expf(-go));
  const float c = sf * cbuf[ci] + si * tanhf(gg);
  const float h = so * tanhf(c);
  cbuf[ci] = c;
  hnext[(size_t)b * H + u] = h;
  out[((size_t)b * T + t) * 2 * H + dir * H + u] = h;
}

// Persistent form of the same recurrence: the launch covers ALL time steps.  A block owns UB hidden units x NB batch
// items of one direction for the whole sequence; its slice of W_hh (H x 4*UB floats, <= 192 KB) is loaded into shared
// memory once, the cell state lives in a register, and the only per-step global traffic is the hidden vector exchange:
// every block publishes its UB units of h_t and the blocks of a direction meet at a counter barrier before reading
// h_t back (L2-coherent loads).  Launched cooperatively (all blocks co-resident, one per SM); a lost barrier traps.
template <int UB>
__global__ void __launch_bounds__(256, 1) lstm_persistent_kernel(const float* __restrict__ pre, const float* __restrict__ whhT,
                                                                 float* __restrict__ hbuf, unsigned int* __restrict__ bar,
                                                                 float* __restrict__ out, int N, int T, int H) {
  constexpr int NB = 256 / UB;
  extern __shared__ float sh[];
  float* sW = sh;                                     // [H][4*UB]   (gate-major inside a row)
  float* sH = sh + (size_t)H * 4 * UB;                // [NB][H]
  const int dir = blockIdx.z, ul = threadIdx.x % UB, bl = threadIdx.x / UB;
  const int u = blockIdx.x * UB + ul, b = blockIdx.y * NB + bl;
  const bool live = u < H && b < N;
  const unsigned int nblk = gridDim.x * gridDim.y;
  for (int i = threadIdx.x; i < H * 4 * UB; i += 256) {
    const int k = i / (4 * UB), r = i - k * 4 * UB, g = r / UB, uu = blockIdx.x * UB + (r - g * UB);
    sW[i] = uu < H ? __ldg(whhT + ((size_t)dir * H + k) * 4 * H + g * H + uu) : 0.f;
  }
  float c = 0.f;
  for (int step = 0; step < T; ++step) {
    const int t = dir ? T - 1 - step : step;
    const float* hprev = hbuf + ((size_t)((step & 1) * 2 + dir) * N) * H;
    float* hnext = hbuf + ((size_t)(((step + 1) & 1) * 2 + dir) * N) * H;
    for (int i = threadIdx.x; i < NB * H; i += 256) {
      const int bb = blockIdx.y * NB + i / H;
      sH[i] = (bb < N && step > 0) ? __ldcg(hprev + (size_t)bb * H + (i % H)) : 0.f;
    }
    __syncthreads();
    if (live) {
      const float* hp = sH + bl * H;
      const float* w = sW + ul;
      float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f;
#pragma unroll 4
      for (int k = 0; k < H; ++k) {
        const float hk = hp[k];
        const float* wk = w + (size_t)k * 4 * UB;
        gi = fmaf(hk, wk[0], gi);
        gf = fmaf(hk, wk[UB], gf);
        gg = fmaf(hk, wk[2 * UB], gg);
        go = fmaf(hk, wk[3 * UB], go);
      }
      const float* p = pre + (((size_t)b * T + t) * 2 + dir) * 4 * H + u;
      gi += __ldg(p); gf += __ldg(p + H); gg += __ldg(p + 2 * H); go += __ldg(p + 3 * H);
      const float si = 1.f / (1.f + expf(-gi)), sf = 1.f / (1.f + expf(-gf)), so = 1.f / (1.f + expf(-go));
      c = sf * c + si * tanhf(gg);
      const float h = so * tanhf(c);
      __stcg(hnext + (size_t)b * H + u, h);
      out[((size_t)b * T + t) * 2 * H + dir * H + u] = h;
    }
    if (step + 1 < T) {       // all blocks of this direction have published h_t before anyone reads it
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&bar[dir], 1u);
        const unsigned int want = (unsigned int)(step + 1) * nblk;
        unsigned int spins = 0;
        while (*reinterpret_cast<volatile unsigned int*>(&bar[dir]) < want) {
          if (++spins > (1u << 27)) __trap();
        }
        __threadfence();
      }
      __syncthreads();
    }
  }
}

// ---- LocalState ----------------------------------------------------------------------------------------------------
// qkc [N, T, 3*D] = (query | key | content) projections, dq [N, T, heads*4] decay logits, out [N, T, D].
// One warp per (item, head, query s): scores over all keys t in shared memory, softmax over t, weighted content sum.
//   dots[t] = k_t . q_s / sqrt(dh) - sum_f (f+1) |t - s| / sqrt(4) * sigmoid(dq[s, f]) / 2;  dots[s] = -100
__global__ void __launch_bounds__(128) local_state_kernel(const float* __restrict__ qkc, const float* __restrict__ dq,
                                                          float* __restrict__ out, int T, int D, int heads) {
  extern __shared__ float sh[];                       // [4 warps][T + dh]
  const int dh = D / heads, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * 4 + warp, h = blockIdx.y, n = blockIdx.z;
  if (s >= T) return;
  float* sc = sh + (size_t)warp * (T + dh);
  float* qs = sc + T;
  const float* base = qkc + (size_t)n * T * 3 * D;
  for (int c = lane; c < dh; c += 32) qs[c] = __ldg(base + (size_t)s * 3 * D + h * dh + c);
  float slope = 0.f;                                  // sum_f (f+1)/2 * sigmoid(dq_f)/2
  for (int f = 0; f < 4; ++f) {
    const float d = __ldg(dq + ((size_t)n * T + s) * heads * 4 + h * 4 + f);
    slope += (float)(f + 1) * 0.5f * (0.5f / (1.0f + expf(-d)));
  }
  __syncwarp();
  const float inv = rsqrtf((float)dh);
  float mx = -INFINITY;
  for (int t = lane; t < T; t += 32) {
    const float* kt = base + (size_t)t * 3 * D + D + h * dh;
    float acc = 0.f;
    for (int c = 0; c < dh; ++c) acc = fmaf(__ldg(kt + c), qs[c], acc);
    float v = acc * inv - slope * fabsf((float)(t - s));
    if (t == s) v = -100.f;
    sc[t] = v;
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int t = lane; t < T; t += 32) {
    const float e = expf(sc[t] - mx);
    sc[t] = e;
    sum += e;
  }
  sum = bd_warp_sum(sum);
  __syncwarp();
  const float r = 1.0f / sum;
  for (int c = lane; c < dh; c += 32) {
    const float* ct = base + 2 * D + h * dh + c;
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc = fmaf(sc[t], __ldg(ct + (size_t)t * 3 * D), acc);
    out[((size_t)n * T + s) * D + h * dh + c] = acc * r;
  }
}

// Same computation with the keys and the content of one (item, head) staged in shared memory once (pitch dh + 1: lanes
// walk the key axis conflict-free) and the block's 8 warps sharing them across all queries: global traffic drops by T.
__global__ void __launch_bounds__(256) local_state_tiled_kernel(const float* __restrict__ qkc, const float* __restrict__ dq,
                                                                float* __restrict__ out, int T, int D, int heads) {
  extern __shared__ float sh[];
  const int dh = D / heads, P = dh + 1, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, n = blockIdx.y;
  float* sK = sh;                                      // [T][P]
  float* sC = sK + (size_t)T * P;                      // [T][P]
  float* sc = sC + (size_t)T * P + (size_t)warp * (T + dh);
  float* qs = sc + T;
  const float* base = qkc + (size_t)n * T * 3 * D;
  for (int i = threadIdx.x; i < T * dh; i += 256) {
    const int t = i / dh, c = i - t * dh;
    sK[t * P + c] = __ldg(base + (size_t)t * 3 * D + D + h * dh + c);
    sC[t * P + c] = __ldg(base + (size_t)t * 3 * D + 2 * D + h * dh + c);
  }
  __syncthreads();
  const float inv = rsqrtf((float)dh);
  for (int s = warp; s < T; s += 8) {
    for (int c = lane; c < dh; c += 32) qs[c] = __ldg(base + (size_t)s * 3 * D + h * dh + c);
    float slope = 0.f;
    for (int f = 0; f < 4; ++f) {
      const float d = __ldg(dq + ((size_t)n * T + s) * heads * 4 + h * 4 + f);
      slope += (float)(f + 1) * 0.5f * (0.5f / (1.0f + expf(-d)));
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int t = lane; t < T; t += 32) {
      const float* kt = sK + t * P;
      float acc = 0.f;
      for (int c = 0; c < dh; ++c) acc = fmaf(kt[c], qs[c], acc);
      float v = acc * inv - slope * fabsf((float)(t - s));
      if (t == s) v = -100.f;
      sc[t] = v;
      mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float e = expf(sc[t] - mx);
      sc[t] = e;
      sum += e;
    }
    sum = bd_warp_sum(sum);
    __syncwarp();
    const float r = 1.0f / sum;
    for (int c = lane; c < dh; c += 32) {
      float acc = 0.f;
      for (int t = 0; t < T; ++t) acc = fmaf(sc[t], sC[t * P + c], acc);
      out[((size_t)n * T + s) * D + h * dh + c] = acc * r;
    }
    __syncwarp();
  }
}

}  // namespace

extern "C" {

int bd_gn_stats(const float* x, double* sums, int B, long long rows, int C, int G, void* stream) {
  BD_REQUIRE(B > 0 && rows > 0 && C > 0 && G > 0 && C % G == 0 && B * G <= 65535, "bd_gn_stats: bad sizes (C=%d G=%d)", C, G);
  long long n = rows * (C / G);
  int gx = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (gx > 64) gx = 64;
  gn_stats_kernel<<<dim3(gx, B * G), 256, 0, (cudaStream_t)stream>>>(x, sums, rows, C, G);
  return bd_check_launch("gn_stats_kernel");
}

int bd_gn_act(const float* x, float* y, const float* mean_rstd, const float* gamma, const float* beta,
              const float* addend, int B, long long rows_in, long long row0, long long rows_out, int C, int G, int act,
              long long y_item_stride, void* stream) {
  BD_REQUIRE(B > 0 && B <= 65535 && rows_in > 0 && row0 >= 0 && rows_out > 0 && row0 + rows_out <= rows_in && C % G == 0 &&
                 (act != BD_ACT_GLU || (C % 2 == 0 && (C / 2) % (C / G) == 0 || G == 1)),
             "bd_gn_act: bad sizes");
  const int Co = act == BD_ACT_GLU ? C / 2 : C;
  gn_act_kernel<<<dim3(bd_cdiv(rows_out * Co, 256), B), 256, 0, (cudaStream_t)stream>>>(x, y, mean_rstd, gamma, beta, addend,
                                                                                        rows_in, row0, rows_out, C, G, act,
                                                                                        y_item_stride);
  return bd_check_launch("gn_act_kernel");
}

int bd_lstm_frame(const float* x, float* frames, int B, long long T, int C, int nframes, int width, int stride, void* stream) {
  BD_REQUIRE(B > 0 && T > 0 && C > 0 && nframes > 0 && width > 0 && stride > 0 && B * nframes <= 65535, "bd_lstm_frame: bad sizes");
  frame_kernel<<<dim3(bd_cdiv((long long)width * C, 256), B * nframes), 256, 0, (cudaStream_t)stream>>>(x, frames, T, C, nframes,
                                                                                                       width, stride);
  return bd_check_launch("frame_kernel");
}

int bd_lstm_unframe_add(const float* frames, const float* skip, float* out, int B, long long T, int C, int nframes, int width,
                        int stride, void* stream) {
  BD_REQUIRE(B > 0 && B <= 65535 && T > 0 && C > 0 && nframes > 0 && width > 0 && stride > 0 &&
                 (long long)(nframes - 1) * stride + width >= T, "bd_lstm_unframe_add: bad sizes");
  unframe_add_kernel<<<dim3(bd_cdiv(T * C, 256), B), 256, 0, (cudaStream_t)stream>>>(frames, skip, out, T, C, nframes, width,
                                                                                   stride);
  return bd_check_launch("unframe_add_kernel");
}

int bd_lstm_bidir(const float* pre, const float* whhT, float* out, float* ws, int N, int T, int H, void* stream) {
  BD_REQUIRE(N > 0 && T > 0 && H > 0 && H <= 1024 && (N + 3) / 4 <= 65535, "bd_lstm_bidir: bad sizes (N=%d T=%d H=%d)", N, T, H);
  const cudaStream_t st = (cudaStream_t)stream;
  float* hbuf = ws;                                    // [2][2][N][H]
  float* cbuf = ws + (size_t)4 * N * H;                // [2][N][H]
  cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)6 * N * H * sizeof(float), st);
  if (e != cudaSuccess) {
    bd_set_error("bd_lstm_bidir: memset: %s", cudaGetErrorString(e));
    return BD_ERR_CUDA;
  }
  // persistent form when a block's slice of W_hh fits in shared memory and the grid fits on the device at once
  const int UB = H <= 192 ? 64 : 32, NB = 256 / UB;
  const size_t smem = ((size_t)H * 4 * UB + (size_t)NB * H) * sizeof(float);
  static const bool stepwise = getenv("BD_LSTM_STEPWISE") != nullptr;
  int dev = 0, sms = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  const int per_chunk = sms / (2 * ((H + UB - 1) / UB));           // batch tiles per launch: grid <= one block per SM
  if (!stepwise && coop && smem <= 227 * 1024 && per_chunk >= 1) {
    auto kern = UB == 64 ? lstm_persistent_kernel<64> : lstm_persistent_kernel<32>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      bd_set_error("bd_lstm_bidir: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return BD_ERR_CUDA;
    }
    for (int n0 = 0; n0 < N; n0 += per_chunk * NB) {
      const int n = N - n0 < per_chunk * NB ? N - n0 : per_chunk * NB;
      const float* pre_c = pre + (size_t)n0 * T * 8 * H;
      float* out_c = out + (size_t)n0 * T * 2 * H;
      float* hb = hbuf;                                   // [2][2][n][H] for this chunk (the buffer holds N >= n)
      unsigned int* bar = reinterpret_cast<unsigned int*>(cbuf);
      if (n0 > 0) {
        e = cudaMemsetAsync(cbuf, 0, 2 * sizeof(unsigned int), st);
        if (e != cudaSuccess) {
          bd_set_error("bd_lstm_bidir: memset: %s", cudaGetErrorString(e));
          return BD_ERR_CUDA;
        }
      }
      int nn = n, TT = T, HH = H;
      void* args[] = {(void*)&pre_c, (void*)&whhT, (void*)&hb, (void*)&bar, (void*)&out_c, (void*)&nn, (void*)&TT, (void*)&HH};
      const dim3 grid((H + UB - 1) / UB, (n + NB - 1) / NB, 2);
      e = cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(256), args, smem, st);
      if (e != cudaSuccess) {
        bd_set_error("bd_lstm_bidir: cooperative launch (grid %d x %d x 2, %zu B smem): %s", grid.x, grid.y, smem,
                     cudaGetErrorString(e));
        return BD_ERR_CUDA;
      }
    }
    return bd_check_launch("lstm_persistent_kernel");
  }
  const dim3 grid((H + 63) / 64, (N + 3) / 4, 2);
  for (int step = 0; step < T; ++step)
    lstm_step_kernel<<<grid, 256, 4 * H * sizeof(float), st>>>(pre, whhT, hbuf, cbuf, out, N, T, H, step);
  return bd_check_launch("lstm_step_kernel");
}

int bd_local_state(const float* qkc, const float* dq, float* out, int N, int T, int D, int heads, void* stream) {
  BD_REQUIRE(N > 0 && N <= 65535 && T > 0 && D > 0 && heads > 0 && heads <= 65535 && D % heads == 0,
             "bd_local_state: bad sizes (T=%d D=%d heads=%d)", T, D, heads);
  const int dh = D / heads;
  const size_t tiled = ((size_t)2 * T * (dh + 1) + (size_t)8 * (T + dh)) * sizeof(float);
  if (tiled <= 200 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(local_state_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled);
    if (e != cudaSuccess) {
      bd_set_error("bd_local_state: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return BD_ERR_CUDA;
    }
    local_state_tiled_kernel<<<dim3(heads, N), 256, tiled, (cudaStream_t)stream>>>(qkc, dq, out, T, D, heads);
    return bd_check_launch("local_state_tiled_kernel");
  }
  const int smem = 4 * (T + D / heads) * (int)sizeof(float);
  BD_REQUIRE(smem <= 48 * 1024, "bd_local_state: T=%d too long for the shared-memory score rows", T);
  local_state_kernel<<<dim3((T + 3) / 4, heads, N), 128, smem, (cudaStream_t)stream>>>(qkc, dq, out, T, D, heads);
  return bd_check_launch("local_state_kernel");
}

}  // extern "C"

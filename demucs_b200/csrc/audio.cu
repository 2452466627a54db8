// Front and back door of the separator on the device (reference demucs/audio.py:143-172,175-265, api.py:265-266):
//   * channel conversion + fractional resampling to the model's rate (convert_audio -> julius.resample_frac): a
//     polyphase windowed-sinc FIR, one output sample per thread, the filter bank [new_sr][2*width + old_sr] read
//     through the read-only cache, input edges replicated as julius pads them;
//   * clip prevention (rescale by the global peak / clamp / tanh) and PCM quantisation fused with the planar ->
//     interleaved transpose, so that stems leave the GPU in wire format (2 bytes per sample instead of 4).
// HBM-bound elementwise / short-FIR kernels; nothing here is on the per-segment hot path.
#include "common.cuh"
#include "../../include/demucs_b200.h"

namespace {

// input channel mix of convert_audio_channels (audio.py:143-166): dst == src: copy; dst == 1: mean of all;
// src == 1: replicate; src > dst: the first dst channels
__device__ __forceinline__ float read_channel(const float* __restrict__ x, int src_ch, int dst_ch, int c, long long len,
                                              long long t) {
  if (src_ch == dst_ch || src_ch > dst_ch && dst_ch != 1) return __ldg(x + (size_t)c * len + t);
  if (src_ch == 1) return __ldg(x + t);
  float s = 0.f;                      // dst_ch == 1: downmix
  for (int k = 0; k < src_ch; ++k) s += __ldg(x + (size_t)k * len + t);
  return s / (float)src_ch;
}

__global__ void convert_channels_kernel(const float* __restrict__ x, float* __restrict__ y, int items, int src_ch,
                                        int dst_ch, long long len) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y % dst_ch, b = blockIdx.y / dst_ch;
  if (t >= len) return;
  y[((size_t)b * dst_ch + c) * len + t] = read_channel(x + (size_t)b * src_ch * len, src_ch, dst_ch, c, len, t);
}

// y[b, c, j*new_sr + i] = sum_k kernel[i][k] * xpad[b, c, j*old_sr + k],  xpad[n] = x[clamp(n - width, 0, Lin-1)]
__global__ void resample_frac_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ kern,
                                     int src_ch, int dst_ch, long long Lin, long long Lout, int old_sr, int new_sr,
                                     int width) {
  const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y % dst_ch, b = blockIdx.y / dst_ch;
  if (o >= Lout) return;
  const long long j = o / new_sr;
  const int i = (int)(o - j * new_sr);
  const int klen = 2 * width + old_sr;
  const float* kr = kern + (size_t)i * klen;
  const float* xb = x + (size_t)b * src_ch * Lin;
  const long long base = j * old_sr - width;
  float acc = 0.f;
  for (int k = 0; k < klen; ++k) {
    long long n = base + k;
    n = n < 0 ? 0 : (n >= Lin ? Lin - 1 : n);
    acc = fmaf(__ldg(kr + k), read_channel(xb, src_ch, dst_ch, c, Lin, n), acc);
  }
  y[((size_t)b * dst_ch + c) * Lout + o] = acc;
}

__global__ void absmax_kernel(const float* __restrict__ x, unsigned int* __restrict__ out, long long n) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));      // non-negative floats order like their bits
}

// x [C, T] planar float -> out [T, C] interleaved: int16 (bits 16), int32 holding a 24-bit value (bits 24), float (bits 32)
__global__ void clip_pcm_kernel(const float* __restrict__ x, void* __restrict__ out, int C, long long T, int mode,
                                const unsigned int* __restrict__ peak_bits, int bits) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float scale = 1.f;
  if (mode == BD_CLIP_RESCALE) scale = 1.0f / fmaxf(1.01f * __uint_as_float(__ldg(peak_bits)), 1.0f);
  for (int c = 0; c < C; ++c) {
    float v = __ldg(x + (size_t)c * T + t);
    if (mode == BD_CLIP_RESCALE) v = v * scale;
    else if (mode == BD_CLIP_CLAMP) v = fminf(fmaxf(v, -0.99f), 0.99f);
    else if (mode == BD_CLIP_TANH) v = tanhf(v);
    const size_t at = (size_t)t * C + c;
    if (bits == 32) {
      reinterpret_cast<float*>(out)[at] = v;
    } else {
      v = fminf(fmaxf(v, -1.f), 1.f);                                    // i16_pcm (audio.py:175-180): clamp, scale,
      if (bits == 16) reinterpret_cast<short*>(out)[at] = (short)(int)(v * 32767.f);      // truncate toward zero
      else reinterpret_cast<int*>(out)[at] = (int)(v * 8388607.f);
    }
  }
}

}  // namespace

extern "C" {

int bd_convert_channels(const float* x, float* y, int items, int src_ch, int dst_ch, long long len, void* stream) {
  BD_REQUIRE(items > 0 && src_ch > 0 && dst_ch > 0 && len > 0 && items * dst_ch <= 65535, "bd_convert_channels: bad sizes");
  BD_REQUIRE(src_ch == dst_ch || dst_ch == 1 || src_ch == 1 || src_ch >= dst_ch,
             "The audio file has less channels than requested but is not mono.");
  convert_channels_kernel<<<dim3(bd_cdiv(len, 256), items * dst_ch), 256, 0, (cudaStream_t)stream>>>(x, y, items, src_ch,
                                                                                                    dst_ch, len);
  return bd_check_launch("convert_channels_kernel");
}

int bd_resample_frac(const float* x, float* y, const float* kernel, int items, int src_ch, int dst_ch, long long Lin,
                     long long Lout, int old_sr, int new_sr, int width, void* stream) {
  BD_REQUIRE(items > 0 && src_ch > 0 && dst_ch > 0 && Lin > 0 && Lout > 0 && old_sr > 0 && new_sr > 0 && width > 0 &&
                 items * dst_ch <= 65535, "bd_resample_frac: bad sizes");
  BD_REQUIRE(src_ch == dst_ch || dst_ch == 1 || src_ch == 1 || src_ch >= dst_ch,
             "The audio file has less channels than requested but is not mono.");
  BD_REQUIRE(Lout <= (Lin * new_sr + old_sr - 1) / old_sr, "bd_resample_frac: Lout beyond ceil(Lin * new_sr / old_sr)");
  resample_frac_kernel<<<dim3(bd_cdiv(Lout, 256), items * dst_ch), 256, 0, (cudaStream_t)stream>>>(
      x, y, kernel, src_ch, dst_ch, Lin, Lout, old_sr, new_sr, width);
  return bd_check_launch("resample_frac_kernel");
}

int bd_absmax(const float* x, float* peak, long long n, void* stream) {
  BD_REQUIRE(n > 0 && x && peak, "bd_absmax: bad arguments");
  cudaError_t e = cudaMemsetAsync(peak, 0, sizeof(float), (cudaStream_t)stream);
  if (e != cudaSuccess) {
    bd_set_error("bd_absmax: memset: %s", cudaGetErrorString(e));
    return BD_ERR_CUDA;
  }
  long long g = (n + 256 * 8 - 1) / (256 * 8);
  if (g > 148 * 8) g = 148 * 8;
  absmax_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<unsigned int*>(peak), n);
  return bd_check_launch("absmax_kernel");
}

int bd_clip_pcm(const float* x, void* out, int channels, long long frames, int mode, const float* peak, int bits,
                void* stream) {
  BD_REQUIRE(channels > 0 && frames > 0, "bd_clip_pcm: bad sizes");
  BD_REQUIRE(bits == 16 || bits == 24 || bits == 32, "bd_clip_pcm: bits must be 16, 24 or 32 (float)");
  BD_REQUIRE(mode >= BD_CLIP_NONE && mode <= BD_CLIP_TANH, "Invalid mode %d", mode);
  BD_REQUIRE(mode != BD_CLIP_RESCALE || peak, "bd_clip_pcm: rescale needs the peak (bd_absmax)");
  clip_pcm_kernel<<<bd_cdiv(frames, 256), 256, 0, (cudaStream_t)stream>>>(x, out, channels, frames, mode,
                                                                          reinterpret_cast<const unsigned int*>(peak), bits);
  return bd_check_launch("clip_pcm_kernel");
}

}  // extern "C"

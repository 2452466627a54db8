// K8: overlap-add of separated segments with the triangular transition weight.
//
// Replaces the accumulation loop of the reference's apply_model split branch
// (demucs/apply.py:271-299): out[..., off:off+sl] += weight[:n] * chunk; sum_weight += weight[:n];
// out /= sum_weight -- plus the centre trim of every chunk (utils.py:38-54), the un-shift and
// averaging of the shift trick (apply.py:253-255) and the bag weighting (apply.py:219-228).
// Gather formulation: one thread owns one output sample and sums the (<= ceil(sl/stride)) segments
// covering it in ascending segment order, so there are no atomics, no zero-fill pass, and the
// summation order equals the reference's.  HBM-bound: reads each segment sample once, writes once.
#include "common.cuh"
#include "../../include/demucs_b200.h"

namespace {

constexpr int OLA_VEC = 4;   // consecutive samples per thread (independent loads in flight)

__global__ void overlap_add_kernel(const float* __restrict__ segs, const float* __restrict__ weight,
                                   float* __restrict__ out, int seg_first, int nseg_local, int nseg, int rows, int valid,
                                   int seg_len, int stride, long long length, long long out_ld, long long out_shift,
                                   long long n_begin, long long n_end, const float* __restrict__ row_alpha, float alpha,
                                   int accumulate) {
  const long long nb = n_begin + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * OLA_VEC;
  const int r = blockIdx.y;
  if (nb >= n_end) return;
  // all OLA_VEC samples of this thread see the same candidate segments except at segment edges, which the
  // per-sample bounds below handle; a sharded caller holds segments [seg_first, seg_first + nseg_local)
  int i_hi = (int)((nb + OLA_VEC - 1) / stride);
  if (i_hi > nseg - 1) i_hi = nseg - 1;
  long long lo = nb - seg_len + 1;
  int i_lo = lo <= 0 ? 0 : (int)((lo + stride - 1) / stride);
  if (i_lo < seg_first) i_lo = seg_first;
  if (i_hi > seg_first + nseg_local - 1) i_hi = seg_first + nseg_local - 1;
  float num[OLA_VEC], den[OLA_VEC];
#pragma unroll
  for (int v = 0; v < OLA_VEC; ++v) num[v] = den[v] = 0.f;
  for (int i = i_lo; i <= i_hi; ++i) {
    const long long off = (long long)i * stride;
    long long rem = length - off;
    const int n_i = rem < seg_len ? (int)rem : seg_len;   // TensorChunk length clip (apply.py:91-94)
    const int lead = (valid - n_i) / 2;                    // centre trim (utils.py:52-53)
    const float* sp = segs + ((size_t)(i - seg_first) * rows + r) * valid + lead;
#pragma unroll
    for (int v = 0; v < OLA_VEC; ++v) {
      const long long k = nb + v - off;
      if (k >= 0 && k < n_i && nb + v < n_end) {
        const float w = __ldg(weight + k);
        num[v] = fmaf(w, __ldg(sp + k), num[v]);
        den[v] += w;
      }
    }
  }
  const float a = alpha * (row_alpha ? __ldg(row_alpha + r) : 1.f);
  float* o = out + (size_t)r * out_ld + (nb - out_shift);
#pragma unroll
  for (int v = 0; v < OLA_VEC; ++v) {
    if (nb + v < n_end) {
      const float val = num[v] / den[v] * a;
      o[v] = accumulate ? o[v] + val : val;
    }
  }
}

}  // namespace

extern "C" int bd_overlap_add(const float* segs, const float* weight, float* out, int seg_first, int nseg_local,
                              int nseg, int rows, int valid, int seg_len, int stride, long long length, long long out_ld,
                              long long out_shift, long long n_begin, long long n_end, const float* row_alpha,
                              float alpha, int accumulate, void* stream) {
  BD_REQUIRE(nseg > 0 && rows > 0 && rows <= 65535 && stride > 0 && seg_len > 0 && (seg_len <= valid || length <= valid),
             "bd_overlap_add: bad sizes (nseg=%d rows=%d seg_len=%d valid=%d stride=%d)", nseg, rows, seg_len, valid, stride);
  BD_REQUIRE((long long)(nseg - 1) * stride < length && (long long)nseg * stride >= length,
             "bd_overlap_add: nseg=%d does not tile length=%lld with stride=%d", nseg, length, stride);
  BD_REQUIRE(out_shift >= 0 && out_shift < length, "bd_overlap_add: out_shift outside the window");
  BD_REQUIRE(seg_first >= 0 && nseg_local > 0 && seg_first + nseg_local <= nseg, "bd_overlap_add: bad segment block");
  if (n_begin < out_shift) n_begin = out_shift;
  if (n_end > length) n_end = length;
  if (n_end <= n_begin) return BD_OK;
  overlap_add_kernel<<<dim3(bd_cdiv(n_end - n_begin, 256 * OLA_VEC), rows), 256, 0, (cudaStream_t)stream>>>(
      segs, weight, out, seg_first, nseg_local, nseg, rows, valid, seg_len, stride, length, out_ld, out_shift, n_begin,
      n_end, row_alpha, alpha, accumulate);
  return bd_check_launch("overlap_add_kernel");
}

// ---- segment gather: the input side of the batcher ------------------------------------------------------------
// batch[(j*B + b), c, t] = track[b, c, start_j + t] (zero outside [0, track_len)), j < nseg_batch, for the segments
// i = seg_first + j of a pass whose window starts at `offset0` in the track: segment i covers window samples
// [i*stride, i*stride + n_i), n_i = min(length - i*stride, seg_len), and is centred in `valid` samples --
// start_j = offset0 + i*stride - (valid - n_i)/2 -- exactly TensorChunk.padded (reference apply.py:108-124): the
// padding is real signal where the track has it, zeros beyond its ends.  One launch per forward batch instead of one
// copy per segment.
namespace {
__global__ void gather_segments_kernel(const float* __restrict__ track, float* __restrict__ batch, int B, int C,
                                       long long track_len, long long offset0, long long length, int seg_first,
                                       int seg_len, int stride, int valid) {
  const int j = blockIdx.z, bc = blockIdx.y;                 // bc = b*C + c
  const long long i = seg_first + j;
  const long long rem = length - i * stride;
  const int n_i = rem < seg_len ? (int)rem : seg_len;
  const long long start = offset0 + i * stride - (valid - n_i) / 2;
  const float* src = track + (size_t)bc * track_len;
  float* dst = batch + ((size_t)j * B * C + bc) * valid;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < valid; t += gridDim.x * blockDim.x) {
    const long long s = start + t;
    dst[t] = (s >= 0 && s < track_len) ? __ldg(src + s) : 0.f;
  }
}
}  // namespace

extern "C" int bd_gather_segments(const float* track, float* batch, int B, int C, long long track_len, long long offset0,
                                  long long length, int seg_first, int nseg_batch, int seg_len, int stride, int valid,
                                  void* stream) {
  BD_REQUIRE(B > 0 && C > 0 && B * C <= 65535 && nseg_batch > 0 && nseg_batch <= 65535 && valid > 0 && seg_len > 0 &&
                 (seg_len <= valid || length <= valid) && stride > 0 && track_len > 0 && length > 0,
             "bd_gather_segments: bad sizes");
  BD_REQUIRE((long long)(seg_first + nseg_batch - 1) * stride < length, "bd_gather_segments: segment beyond the window");
  int gx = bd_cdiv(valid, 256 * 4);
  if (gx > 1024) gx = 1024;
  gather_segments_kernel<<<dim3(gx, B * C, nseg_batch), 256, 0, (cudaStream_t)stream>>>(
      track, batch, B, C, track_len, offset0, length, seg_first, seg_len, stride, valid);
  return bd_check_launch("gather_segments_kernel");
}

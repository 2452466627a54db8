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

__global__ void overlap_add_kernel(const float* __restrict__ segs, const float* __restrict__ weight,
                                   float* __restrict__ out, int seg_first, int nseg_local, int nseg, int rows, int valid,
                                   int seg_len, int stride, long long length, long long out_ld, long long out_shift,
                                   long long n_begin, long long n_end, const float* __restrict__ row_alpha, float alpha,
                                   int accumulate) {
  const long long n = n_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (n >= n_end) return;
  int i_hi = (int)(n / stride);
  if (i_hi > nseg - 1) i_hi = nseg - 1;
  long long lo = n - seg_len + 1;
  int i_lo = lo <= 0 ? 0 : (int)((lo + stride - 1) / stride);
  // a sharded caller holds segments [seg_first, seg_first + nseg_local); it must own sample n entirely
  if (i_lo < seg_first) i_lo = seg_first;
  if (i_hi > seg_first + nseg_local - 1) i_hi = seg_first + nseg_local - 1;
  float num = 0.f, den = 0.f;
  for (int i = i_lo; i <= i_hi; ++i) {
    const long long off = (long long)i * stride;
    const int k = (int)(n - off);
    long long rem = length - off;
    const int n_i = rem < seg_len ? (int)rem : seg_len;   // TensorChunk length clip (apply.py:91-94)
    if (k >= n_i) continue;
    const int lead = (valid - n_i) / 2;                    // centre trim (utils.py:52-53)
    const float w = __ldg(weight + k);
    num = fmaf(w, __ldg(segs + ((size_t)(i - seg_first) * rows + r) * valid + lead + k), num);
    den += w;
  }
  float v = num / den;
  v *= alpha * (row_alpha ? __ldg(row_alpha + r) : 1.f);
  float* o = out + (size_t)r * out_ld + (n - out_shift);
  *o = accumulate ? *o + v : v;
}

}  // namespace

extern "C" int bd_overlap_add(const float* segs, const float* weight, float* out, int seg_first, int nseg_local,
                              int nseg, int rows, int valid, int seg_len, int stride, long long length, long long out_ld,
                              long long out_shift, long long n_begin, long long n_end, const float* row_alpha,
                              float alpha, int accumulate, void* stream) {
  BD_REQUIRE(nseg > 0 && rows > 0 && rows <= 65535 && stride > 0 && seg_len > 0 && seg_len <= valid,
             "bd_overlap_add: bad sizes (nseg=%d rows=%d seg_len=%d valid=%d stride=%d)", nseg, rows, seg_len, valid, stride);
  BD_REQUIRE((long long)(nseg - 1) * stride < length && (long long)nseg * stride >= length,
             "bd_overlap_add: nseg=%d does not tile length=%lld with stride=%d", nseg, length, stride);
  BD_REQUIRE(out_shift >= 0 && out_shift < length, "bd_overlap_add: out_shift outside the window");
  BD_REQUIRE(seg_first >= 0 && nseg_local > 0 && seg_first + nseg_local <= nseg, "bd_overlap_add: bad segment block");
  if (n_begin < out_shift) n_begin = out_shift;
  if (n_end > length) n_end = length;
  if (n_end <= n_begin) return BD_OK;
  overlap_add_kernel<<<dim3(bd_cdiv(n_end - n_begin, 256), rows), 256, 0, (cudaStream_t)stream>>>(
      segs, weight, out, seg_first, nseg_local, nseg, rows, valid, seg_len, stride, length, out_ld, out_shift, n_begin,
      n_end, row_alpha, alpha, accumulate);
  return bd_check_launch("overlap_add_kernel");
}

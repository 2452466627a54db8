// K6 (tensor-core arm): fused attention softmax(Q K^T / 8) V on tcgen05 / TMEM / TMA, head_dim 64.
//
// One CTA = 128 queries of one (item, head); it walks the keys in tiles of 128.
//   S = Q K^T      : UMMA 128x128x8 (tf32), Q and K tiles are K-major SWIZZLE_128B boxes straight from TMA
//   P = softmax    : 4 warps, one query row per thread (TMEM lane = row, so row max / sum need no shuffles),
//                    exp2 domain, running max / sum in registers; P is written back to TENSOR MEMORY
//   O += P V       : UMMA 128x64x8 with the A operand read from TMEM (P never touches shared memory) and V^T
//                    tiles as the K-major B operand.  TF32 operands that are MN-major need the special
//                    SWIZZLE_128B_BASE32B layout; instead a small tiled-transpose kernel writes V^T
//                    [B, H*64, Tk] once per call (2 x 4 bytes per V element, ~2% of the attention time)
// S is double-buffered in TMEM so that Q K^T of tile j+1 runs on the tensor core while the softmax warps work
// on tile j; K and V tiles are double-buffered in shared memory.  TMEM: S0 | S1 | P | O = 128+128+128+64 columns.
// Token counts are not multiples of 128 (1344 = 10.5 tiles): TMA zero-fills rows past the item, the softmax
// masks them.  Replaces the SDPA core of nn.MultiheadAttention (reference transformer.py:365,506).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/demucs_b200.h"

namespace {

constexpr int AT_THREADS = 192;
constexpr int TQ = 128, TK = 128, HD = 64;
constexpr int SUB = TQ * 32 * 4;                 // one 128-row x 32-float swizzled box = 16 KB
constexpr int VSUB = HD * 32 * 4;                // one 64-row x 32-float box of V^T = 8 KB
constexpr int kQBytes = 2 * SUB, kKVBytes = 2 * SUB;
constexpr int AT_SMEM = kQBytes + 2 * kKVBytes + 2 * kKVBytes + 1024 + 256;
constexpr uint32_t kSpinLimit = 1u << 26;
constexpr uint32_t COL_S0 = 0, COL_S1 = 128, COL_P = 256, COL_O = 384;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// K-major SWIZZLE_128B operand (rows of 32 floats): SBO = 1024 B between 8-row groups
__device__ __forceinline__ uint64_t desc_kmajor(const void* smem) {
  return (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_128B operand: a 128-byte row holds 32 consecutive MN (= head-dim) elements of one K (= key)
// index; 8 keys form the 1024-byte swizzle atom (SBO), the next 32 head-dim elements start LBO bytes further.
__device__ __forceinline__ uint64_t desc_mnmajor(const void* smem, uint32_t lbo_bytes) {
  return (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(AT_THREADS) attention_tc_kernel(const __grid_constant__ CUtensorMap map_q,
                                                                  const __grid_constant__ CUtensorMap map_k,
                                                                  const __grid_constant__ CUtensorMap map_v,
                                                                  float* __restrict__ o, int Tq, int Tk, int ldo,
                                                                  int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;                    // 2 stages
  uint8_t* sV = sK + 2 * kKVBytes;               // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * kKVBytes);
  uint64_t* q_full = bars;                       // [1]
  uint64_t* k_full = bars + 1;                   // [2]
  uint64_t* v_full = bars + 3;                   // [2]
  uint64_t* k_empty = bars + 5;                  // [2]
  uint64_t* v_empty = bars + 7;                  // [2]
  uint64_t* s_full = bars + 9;                   // [2]
  uint64_t* p_full = bars + 11;                  // [1] 128 arrivals
  uint64_t* o_done = bars + 12;                  // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int ntiles = (Tk + TK - 1) / TK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 11; ++i) mbar_init(&bars[i], 1);
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(q_full, kQBytes);
      tma_load_3d(&map_q, q_full, sQ, h * HD, q0, b);
      tma_load_3d(&map_q, q_full, sQ + SUB, h * HD + 32, q0, b);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], kKVBytes);
        tma_load_3d(&map_k, &k_full[s], sK + s * kKVBytes, h * HD, j * TK, b);
        tma_load_3d(&map_k, &k_full[s], sK + s * kKVBytes + SUB, h * HD + 32, j * TK, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], kKVBytes);
#pragma unroll
        for (int i = 0; i < 4; ++i)              // V^T tile: 64 head-dim rows x 128 keys = 4 boxes of 32 keys
          tma_load_3d(&map_v, &v_full[s], sV + s * kKVBytes + i * VSUB, j * TK + 32 * i, h * HD, b);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc_qk = idesc_tf32(TQ, TK, 0);
      constexpr uint32_t idesc_pv = idesc_tf32(TQ, HD, 0);
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem + (s ? COL_S1 : COL_S0);
#pragma unroll
        for (int kk = 0; kk < HD / 8; ++kk) {   // 8 k-steps: two 32-float boxes x 4
          const uint64_t a = desc_kmajor(sQ + (kk >> 2) * SUB) + 2 * (kk & 3);
          const uint64_t bq = desc_kmajor(sK + s * kKVBytes + (kk >> 2) * SUB) + 2 * (kk & 3);
          umma_ss(d, a, bq, idesc_qk, kk != 0);
        }
        tc_commit(&s_full[s]);
        tc_commit(&k_empty[s]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_qk(j + 1);   // overlaps the softmax of tile j
        const int s = j & 1;
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        mbar_wait(&v_full[s], (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < TK / 8; ++kk) {   // 16 k-steps of 8 keys
          const uint64_t bv = desc_kmajor(sV + s * kKVBytes + (kk >> 2) * VSUB) + 2 * (kk & 3);
          umma_ts(tmem + COL_O, tmem + (dbg == 4 ? COL_S0 : COL_P) + kk * 8, bv, idesc_pv, (j | kk) != 0);
        }
        tc_commit(o_done);
        tc_commit(&v_empty[s]);
      }
    }
  } else {
    // ===== softmax warps: thread = query row =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float sl2 = 0.125f * 1.44269504088896340736f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1;
      const int kv_valid = min(TK, Tk - j * TK);
      mbar_wait(&s_full[s], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem + lane_addr + (s ? COL_S1 : COL_S0);
      // pass 1: row maximum
      float mx = -INFINITY;
      for (int c0 = 0; c0 < TK; c0 += 32) {
        if (c0 >= kv_valid) break;
        uint32_t v[32];
        tmem_ld32(s_addr + c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, mx * sl2);
      const float corr = exp2f(m_run - m_new);
      // the previous P V product must have drained before P is overwritten / O is rescaled
      if (j > 0) {
        mbar_wait(o_done, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, corr != 1.0f)) {
          for (int c0 = 0; c0 < HD; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + lane_addr + COL_O + c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * corr);
            tmem_st32(tmem + lane_addr + COL_O + c0, v);
          }
        }
      }
      // pass 2: probabilities -> TMEM
      float rs = 0.f;
      for (int c0 = 0; c0 < TK; c0 += 32) {
        uint32_t v[32];
        if (c0 < kv_valid) {
          tmem_ld32(s_addr + c0, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float p = (c0 + i < kv_valid) ? exp2f(fmaf(__uint_as_float(v[i]), sl2, -m_new)) : 0.f;
            if (dbg >= 2) p = (c0 + i < kv_valid) ? 1.f : 0.f;
            rs += p;
            v[i] = __float_as_uint(p);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        tmem_st32(tmem + lane_addr + COL_P + c0, v);
      }
      tmem_st_wait();
      l_run = l_run * corr + rs;
      m_run = m_new;
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ===== output: O / l =====
    mbar_wait(o_done, (ntiles - 1) & 1);
    tc_fence_after();
    const int r = q0 + row;
    const float inv = dbg ? 1.0f : 1.0f / l_run;
    for (int c0 = 0; c0 < HD; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + lane_addr + (dbg == 1 ? COL_S0 : dbg == 3 ? COL_P : COL_O) + c0, v);
      if (r < Tq) {
        float4* dst = reinterpret_cast<float4*>(o + ((size_t)b * Tq + r) * ldo + h * HD + c0);
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          dst[i >> 2] = make_float4(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv,
                                    __uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_map3(CUtensorMap* map, const float* base, int cols, int T, int B, int ld, int box_rows = 128,
               long long item_stride = 0) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return false;
    enc = (EncodeTiledFn)p;
  }
  cuuint64_t dim[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t str[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(item_stride ? item_stride : (long long)T * ld) * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// v [B, Tk, ldv] (columns 0 .. D-1) -> vt [B, D, Tkp], 32x32 tiles through shared memory
__global__ void transpose_v_kernel(const float* __restrict__ v, float* __restrict__ vt, int Tk, int D, int ldv, int Tkp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i;
    tile[i][tx] = t < Tk ? __ldg(v + ((size_t)b * Tk + t) * ldv + c0 + tx) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + tx;
    if (t < Tkp) vt[((size_t)b * D + c0 + i) * Tkp + t] = tile[tx][i];
  }
}

}  // namespace

int bd_attention_tc(const float* q, const float* k, const float* v, float* o, int B, int H, int Tq, int Tk, int ldq,
                    int ldk, int ldv, int ldo, float* ws, void* stream) {
  BD_REQUIRE(B > 0 && H > 0 && Tq > 0 && Tk > 0, "bd_attention: bad sizes");
  BD_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0, "bd_attention: leading dims must be multiples of 4");
  BD_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o | (uintptr_t)ws) & 15) == 0, "bd_attention: unaligned tensor");
  BD_REQUIRE(ws != nullptr, "bd_attention: the tensor-core arm needs a workspace of B*H*64*ceil4(Tk) floats");
  const int D = H * HD, Tkp = (Tk + 3) & ~3;
  transpose_v_kernel<<<dim3((Tkp + 31) / 32, D / 32, B), 256, 0, (cudaStream_t)stream>>>(v, ws, Tk, D, ldv, Tkp);
  if (bd_check_launch("transpose_v_kernel") != BD_OK) return BD_ERR_CUDA;
  alignas(64) CUtensorMap mq, mk, mv;
  if (!make_map3(&mq, q, D, Tq, B, ldq) || !make_map3(&mk, k, D, Tk, B, ldk) ||
      !make_map3(&mv, ws, Tk, D, B, Tkp, HD, (long long)D * Tkp)) {
    bd_set_error("bd_attention_tc: cuTensorMapEncodeTiled failed");
    return BD_ERR_CUDA;
  }
  cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
  if (e != cudaSuccess) {
    bd_set_error("bd_attention_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return BD_ERR_CUDA;
  }
  dim3 grid((Tq + TQ - 1) / TQ, H, B);
  const char* dbg_env = getenv("BD_ATTN_DEBUG");
  attention_tc_kernel<<<grid, AT_THREADS, AT_SMEM, (cudaStream_t)stream>>>(mq, mk, mv, o, Tq, Tk, ldo,
                                                                           dbg_env ? atoi(dbg_env) : 0);
  return bd_check_launch("attention_tc_kernel");
}

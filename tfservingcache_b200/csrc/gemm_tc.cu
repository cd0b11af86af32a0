// tcgen05 GEMM for the graph executor (conv-as-GEMM, transformer dense layers), sm_100a only:
//
//   C[M,N] = act(A[M,K] (row stride lda) * B[K,N] + bias[N] (+ R[M,N])),  fp32 in / out, 3xTF32 split (fp32-accurate)
//
// Activations are the MMA "A" operand (M = 128 rows per CTA), weights [K,N] row-major the "B" operand:
//   * TMA lands the A tile [128 m][32 k] (K-major, SWIZZLE_128B) and the B tile [32 k][BN n] (MN-major tf32 ->
//     SWIZZLE_128B_ATOM_32B, 3-D map {32 n, K, N/32}) in a 4-stage shared-memory ring.
//   * converter warps 2..5 move the A tile into TMEM (lane = row m, column = k) as A_hi (raw fp32 bits; kind::tf32
//     ignores the low 13 mantissa bits) and A_lo = A - trunc_tf32(A); warps 6..9 write B_lo = B - trunc_tf32(B) next
//     to B_hi in the stage, so B' = [B_hi | B_lo] is one MN-major operand of N = 2*BN columns.
//   * per 8-wide k step: MMA1 D[:, 0:2BN] += A_hi . B'  (hh | hl), MMA2 D[:, BN:2BN] += A_lo . B_hi (lh): the large
//     term and the small corrections have separate TMEM accumulators (the tensor core's fp32 accumulate truncates).
//   * epilogue: tcgen05.ld, C = hh + small + bias (+ residual), ReLU / GELU(erf) / tanh, row-contiguous stores.
// One CTA per 128 x BN output tile over the whole K (no split-K): conv / transformer GEMMs have M in the hundreds to
// tens of thousands. Same role split as dense_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, 8 converter/epilogue warps.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <functional>
#include <mutex>
#include <unordered_map>

#include "kernels.h"
#include "tc_ptx.cuh"

namespace tfsc {

extern std::atomic<int64_t> g_launches_nn;

namespace gt {
constexpr int BM = 128, BK = 32;
constexpr int A_BYTES = BM * BK * 4;  // 16 KB
constexpr int STAGES = 4;
constexpr int THREADS = 320;
constexpr uint32_t kAopCol = 256;     // TMEM: D at [0, 2*BN), A operand staging at [256 + cb*64, +64): hi 32 | lo 32
}  // namespace gt

template <int BN>
struct GtSmem {
  static constexpr int SLABS = BN / 32;                 // 32-column slabs of the B tile
  static constexpr int KG_BYTES = 2 * SLABS * 512;      // one 4-row k group: hi slabs then lo slabs, 512 B each
  static constexpr int B_BYTES = (gt::BK / 4) * KG_BYTES;  // 8 k groups: 2*BN*32*4 bytes
  static constexpr int STAGE_BYTES = gt::A_BYTES + B_BYTES;
  static constexpr int TOTAL = gt::STAGES * STAGE_BYTES + 256 + 1024;
};

__device__ __forceinline__ float gelu_erf_tc(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <int BN>
__global__ void __launch_bounds__(gt::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
               const float* __restrict__ bias, const float* __restrict__ R, float* __restrict__ C, int M, int N, int K, int act) {
  using S = GtSmem<BN>;
  constexpr int NS = gt::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * S::STAGE_BYTES);
  uint64_t* full = bars;            // [NS] TMA landed A and B_hi
  uint64_t* conv = bars + NS;       // [NS] converters published A (TMEM) and B_lo (smem)
  uint64_t* empty = bars + 2 * NS;  // [NS] MMAs finished reading the stage
  uint64_t* cempty = bars + 3 * NS; // [2]  MMAs finished reading TMEM A buffer cb
  uint64_t* accum_full = cempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * gt::BM, n0 = blockIdx.x * BN;
  const int n_kblocks = (K + gt::BK - 1) / gt::BK;
  constexpr int TMEM_COLS = 512;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < NS; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&conv[s], 8);
        mbar_init(&empty[s], 1);
      }
      mbar_init(&cempty[0], 1);
      mbar_init(&cempty[1], 1);
      mbar_init(accum_full, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&bmap) : "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    for (int kb = 0; kb < n_kblocks; ++kb) {
      const int s = kb % NS, it = kb / NS;
      if (lane == 0) {
        if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);
        uint8_t* stage = smem + s * S::STAGE_BYTES;
        mbar_expect_tx(&full[s], gt::A_BYTES + BN * gt::BK * 4);
        tma_load_2d(stage, &amap, &full[s], kb * gt::BK, m0);            // A: [128 m][32 k], 128 B rows, SW128
#pragma unroll
        for (int g = 0; g < gt::BK / 4; ++g)                              // B_hi: k group g -> slabs [0, SLABS) of the group
          tma_load_3d(stage + gt::A_BYTES + g * S::KG_BYTES, &bmap, &full[s], 0, kb * gt::BK + g * 4, n0 / 32);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc1 = make_idesc_ts_b(2 * BN, 1);  // A_hi (TMEM) x [B_hi | B_lo] (MN-major)
    constexpr uint32_t idesc2 = make_idesc_ts_b(BN, 1);      // A_lo (TMEM) x  B_hi
    for (int kb = 0; kb < n_kblocks; ++kb) {
      const int s = kb % NS, it = kb / NS, cb = kb & 1;
      if (lane == 0) {
        mbar_wait(&conv[s], it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t bsm = smem_u32(smem + s * S::STAGE_BYTES + gt::A_BYTES);
        const uint32_t ahi = tmem_base + gt::kAopCol + (uint32_t)(cb * 64);
        const uint32_t alo = ahi + 32;
#pragma unroll
        for (int k8 = 0; k8 < gt::BK / 8; ++k8) {
          // B' (MN-major tf32, SWIZZLE_128B_BASE32B): atoms of 4 k rows x 128 B; slabs LBO = 512 B apart, consecutive
          // 4-row k groups SBO = KG_BYTES apart; one MMA (K = 8) spans two k groups
          const uint64_t b = make_desc(bsm + k8 * 2 * S::KG_BYTES, 512, S::KG_BYTES, 1);
          umma_tf32_ts(tmem_base, ahi + k8 * 8, b, idesc1, (kb | k8) ? 1u : 0u);
          umma_tf32_ts(tmem_base + BN, alo + k8 * 8, b, idesc2, 1u);
        }
        umma_commit(&empty[s]);
        umma_commit(&cempty[cb]);
      }
      __syncwarp();
    }
    if (lane == 0) umma_commit(accum_full);
    __syncwarp();
  } else {
    const int ct = threadIdx.x - 64;   // 0..255
    const int q = warp & 3;            // TMEM lane quarter of this warp
    if (warp < 6) {
      // ===================== A converters (warps 2..5): smem A tile -> TMEM A_hi / A_lo =====================
      const int m = q * 32 + lane;     // row of the tile = TMEM lane
      for (int kb = 0; kb < n_kblocks; ++kb) {
        const int s = kb % NS, it = kb / NS, cb = kb & 1, cit = kb >> 1;
        if (cit > 0) mbar_wait(&cempty[cb], (cit - 1) & 1);
        mbar_wait(&full[s], it & 1);
        const uint32_t arow = smem_u32(smem + s * S::STAGE_BYTES) + (uint32_t)(m * 128);
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {  // 16-byte chunk c of row m sits at chunk (c ^ (m & 7)) (128-byte swizzle)
          const float4 v = lds_f4(arow + (uint32_t)(((c ^ (m & 7)) << 4)));
          hi[4 * c + 0] = __float_as_uint(v.x); lo[4 * c + 0] = __float_as_uint(tf32_lo(v.x));
          hi[4 * c + 1] = __float_as_uint(v.y); lo[4 * c + 1] = __float_as_uint(tf32_lo(v.y));
          hi[4 * c + 2] = __float_as_uint(v.z); lo[4 * c + 2] = __float_as_uint(tf32_lo(v.z));
          hi[4 * c + 3] = __float_as_uint(v.w); lo[4 * c + 3] = __float_as_uint(tf32_lo(v.w));
        }
        const uint32_t aop = tmem_base + ((uint32_t)(q * 32) << 16) + gt::kAopCol + (uint32_t)(cb * 64);
        tmem_st32(aop, hi);
        tmem_st32(aop + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv[s]);
      }
    } else {
      // ===================== B converters (warps 6..9): B_lo = B - trunc_tf32(B), elementwise in place ==========
      const int bt = ct - 128;         // 0..127
      for (int kb = 0; kb < n_kblocks; ++kb) {
        const int s = kb % NS, it = kb / NS;
        mbar_wait(&full[s], it & 1);   // the stage itself is free: the producer waited on empty[s] before refilling
        const uint32_t bsm = smem_u32(smem + s * S::STAGE_BYTES + gt::A_BYTES);
        // hi part of k group g: bytes [g*KG, g*KG + SLABS*512); lo part right behind it
        constexpr int F4_PER_GROUP = S::SLABS * 512 / 16;
        for (int idx = bt; idx < (gt::BK / 4) * F4_PER_GROUP; idx += 128) {
          const int g = idx / F4_PER_GROUP, o = idx - g * F4_PER_GROUP;
          const uint32_t src = bsm + (uint32_t)(g * S::KG_BYTES + o * 16);
          const float4 v = lds_f4(src);
          sts_f4(src + S::SLABS * 512, make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w)));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv[s]);
      }
    }
    // ===================== epilogue: all 8 warps, 4 lane quarters x 2 column halves =====================
    mbar_wait(accum_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int half = (warp - 2) >> 2;                  // warps 2..5 -> columns [0, BN/2), warps 6..9 -> [BN/2, BN)
    const int gm = m0 + q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
    for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 16) {
      float hh[16], sm[16];
      tmem_ld16(taddr + c, hh);
      tmem_ld16(taddr + BN + c, sm);
      if (gm < M) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int gn = n0 + c + j;
          if (gn >= N) continue;
          float4 v = make_float4(hh[j] + sm[j], hh[j + 1] + sm[j + 1], hh[j + 2] + sm[j + 2], hh[j + 3] + sm[j + 3]);
          if (bias) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + gn));
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          }
          if (R) {
            const float4 rv = __ldg(reinterpret_cast<const float4*>(R + (size_t)gm * N + gn));
            v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
          }
          if (act == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          else if (act == 2) { v.x = gelu_erf_tc(v.x); v.y = gelu_erf_tc(v.y); v.z = gelu_erf_tc(v.z); v.w = gelu_erf_tc(v.w); }
          else if (act == 3) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
          *reinterpret_cast<float4*>(C + (size_t)gm * N + gn) = v;
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

// --------------------------------------------------------------------------------- host side ----
struct GKey {
  const void* p;
  int64_t a, b, c;
  bool operator==(const GKey& o) const { return p == o.p && a == o.a && b == o.b && c == o.c; }
};
struct GKeyHash {
  size_t operator()(const GKey& k) const {
    return std::hash<const void*>()(k.p) ^ ((size_t)k.a * 1315423911u) ^ ((size_t)k.b << 21) ^ ((size_t)k.c << 42);
  }
};

static bool cached_map(const GKey& key, const std::function<bool(CUtensorMap*)>& make, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<GKey, CUtensorMap, GKeyHash> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return true;
  }
  CUtensorMap m;
  if (!make(&m)) return false;
  if (cache.size() > 8192) cache.clear();
  cache[key] = m;
  *out = m;
  return true;
}

bool gemm_tc_supported(const float* A, const float* B, const float* bias, const float* R, const float* C, int M, int N, int K,
                       int lda) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return M >= 64 && N >= 64 && N % 32 == 0 && K >= 32 && lda % 4 == 0 && lda >= K && al16(A) && al16(B) && al16(C) &&
         (!bias || al16(bias)) && (!R || al16(R)) && tc_encode_fn() != nullptr;
}

template <int BN>
static cudaError_t launch_gt(const CUtensorMap& am, const CUtensorMap& bm, const float* bias, const float* R, float* C, int M,
                             int N, int K, int act, cudaStream_t s) {
  static bool attr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GtSmem<BN>::TOTAL);
    if (e != cudaSuccess) return e;
    attr[dev & 63] = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + gt::BM - 1) / gt::BM);
  gemm_tc_kernel<BN><<<grid, gt::THREADS, GtSmem<BN>::TOTAL, s>>>(am, bm, bias, R, C, M, N, K, act);
  g_launches_nn++;
  return cudaGetLastError();
}

cudaError_t launch_gemm_tc(const float* A, const float* B, const float* bias, const float* R, float* C, int M, int N, int K,
                           int lda, int act, cudaStream_t s) {
  EncodeTiledFn enc = tc_encode_fn();
  if (!enc) return cudaErrorNotSupported;
  const int BN = (N % 128 == 0 || N > 128) ? 128 : 64;
  CUtensorMap am, bm;
  if (!cached_map({A, M, K, lda}, [&](CUtensorMap* m) {
        const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
        const cuuint64_t gstride[1] = {(cuuint64_t)lda * 4};
        const cuuint32_t box[2] = {(cuuint32_t)gt::BK, (cuuint32_t)gt::BM};
        const cuuint32_t estr[2] = {1, 1};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      }, &am))
    return cudaErrorInvalidValue;
  if (!cached_map({B, K, N, BN}, [&](CUtensorMap* m) {
        const cuuint64_t gdim[3] = {32, (cuuint64_t)K, (cuuint64_t)(N / 32)};
        const cuuint64_t gstride[2] = {(cuuint64_t)N * 4, 128};
        const cuuint32_t box[3] = {32, 4, (cuuint32_t)(BN / 32)};
        const cuuint32_t estr[3] = {1, 1, 1};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(B), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      }, &bm))
    return cudaErrorInvalidValue;
  return BN == 128 ? launch_gt<128>(am, bm, bias, R, C, M, N, K, act, s) : launch_gt<64>(am, bm, bias, R, C, M, N, K, act, s);
}

}  // namespace tfsc

// tc_probe.cu -- unit probe of the tensor-core building block of the learned-lifting kernel (K2-TC):
//   D[128][64] (fp32, TMEM) = A[128][K] (TMEM-resident, tf32) x B[K][64] (shared memory, MN-major,
//   SWIZZLE_128B atoms of 8 k-rows x 32 columns), K in blocks of 8, either plain TF32 (1 MMA per
//   block) or the 3xTF32 split (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo) that keeps fp32-level accuracy.
// No reference counterpart: it exists so the operand layouts / descriptors of lift_tc.cu are pinned by
// a GPU test against a plain fp32 matmul (tests/test_gpu_tc_probe.py) and so the MMA issue rate of
// this shape can be measured (ll_tc_tf32_probe returns SM cycles per chain).
#include <stdint.h>

#include "../ll_common.cuh"
#include "../../../include/ll_probe.h"
#include "../tc_ptx.cuh"

namespace ll {

constexpr int TP_MAXKB = 20;                 // k-blocks: 2 (hi, lo) x 8 x 20 = 320 TMEM columns
constexpr int TP_N = 64;
constexpr int TP_ATOM = 1024;                // bytes: 8 k-rows x 128 B
constexpr int TP_SMEM = 1024 + 2 * TP_MAXKB * 2 * TP_ATOM + 64;

__global__ void __launch_bounds__(128, 1)
tc_tf32_probe_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ D, int kblocks,
                     int split, int reps, long long* __restrict__ cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bar = base + 2 * TP_MAXKB * 2 * TP_ATOM;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + 2 * TP_MAXKB * 2 * TP_ATOM + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = kblocks * 8;

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // A -> TMEM: row m = thread; columns [0,K) hi, [K,2K) lo
  {
    const float* ar = A + (long long)tid * K;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int kb = 0; kb < kblocks; ++kb) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = ar[kb * 8 + j];
        const float h = tf32_rna(a);
        hi[j] = __float_as_uint(h);
        lo[j] = __float_as_uint(tf32_rna(a - h));
      }
      tmem_st8(trow + kb * 8, hi);
      tmem_st8(trow + K + kb * 8, lo);
    }
    tmem_wait_st();
  }
  // B -> smem atoms [term][kb][na]: element (k, n) at k*128 + (((n%32)/4) ^ k)*16 + (n%4)*4
  for (int e = tid; e < K * TP_N; e += 128) {
    const int k = e / TP_N, n = e % TP_N;
    const float b = Bm[e];
    const float h = tf32_rna(b);
    const float l = tf32_rna(b - h);
    const int kb = k >> 3, kr = k & 7, na = n >> 5, p = n & 31;
    uint32_t off = (uint32_t)((kb * 2 + na) * TP_ATOM + kr * 128 + (((p >> 2) ^ kr) << 4) + ((p & 3) << 2));
    if (split & 4) {   // MN-major, SWIZZLE_128B_BASE32B: atom = 4 k-rows x 128 B, 32-byte chunks XOR k-row
      off = (uint32_t)((kb * 2 + na) * TP_ATOM + (kr >> 2) * 512 + (kr & 3) * 128 + ((((p >> 3) ^ (kr & 3))) << 5) + ((p & 7) << 2));
    }
    if (split & 2) {   // debug variant: K-major B, rows = n (128 B = 32 k), 8-row groups of 1024 B
      const int kg = k >> 5, kk = k & 31;
      off = (uint32_t)(kg * 8192 + (n >> 3) * 1024 + (n & 7) * 128 + ((((kk >> 2) ^ (n & 7))) << 4) + ((kk & 3) << 2));
    }
    *reinterpret_cast<float*>(gen + off) = h;
    *reinterpret_cast<float*>(gen + TP_MAXKB * 2 * TP_ATOM + off) = l;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {   // warp-uniform issue loop, one elected lane per instruction
    // D fp32, A/B tf32, A K-major (TMEM), B MN-major, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((split & 2) ? 0u : (1u << 16)) | ((uint32_t)(TP_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t d_tmem = tmem + 384;
    // descriptors advance by adding to the start-address field (16-byte units): cheap single-thread issue
    uint64_t bh0 = umma_desc_mn_sw128(base, TP_ATOM), bl0 = umma_desc_mn_sw128(base + TP_MAXKB * 2 * TP_ATOM, TP_ATOM);
    uint32_t dstep = (2 * TP_ATOM) >> 4;
    if (split & 4) {
      bh0 = umma_desc_mn_sw128_32b(base, TP_ATOM, 512);
      bl0 = umma_desc_mn_sw128_32b(base + TP_MAXKB * 2 * TP_ATOM, TP_ATOM, 512);
    }
    if (split & 2) {
      bh0 = umma_desc_sw128(base);
      bl0 = umma_desc_sw128(base + TP_MAXKB * 2 * TP_ATOM);
    }
    const long long t0 = clock64();
    if ((split & 8) && kblocks == TP_MAXKB) {
      // straight-line issue: every descriptor is the base plus an immediate
      for (int r = 0; r < reps; ++r) {
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < TP_MAXKB; ++kb) {
            const uint64_t bh = bh0 + (uint32_t)(kb * ((2 * TP_ATOM) >> 4)), bl = bl0 + (uint32_t)(kb * ((2 * TP_ATOM) >> 4));
            tc_mma_tf32_ts(d_tmem, tmem + kb * 8, bh, idesc, (uint32_t)(kb != 0));
            tc_mma_tf32_ts(d_tmem, tmem + TP_MAXKB * 8 + kb * 8, bh, idesc, 1u);
            tc_mma_tf32_ts(d_tmem, tmem + kb * 8, bl, idesc, 1u);
          }
        }
        __syncwarp();
      }
    } else
    for (int r = 0; r < reps; ++r) {
#pragma unroll 4
      for (int kb = 0; kb < kblocks; ++kb) {
        const uint32_t adv = (split & 2) ? (uint32_t)((kb >> 2) * (8192 >> 4) + 2 * (kb & 3)) : kb * dstep;
        const uint64_t bh = bh0 + adv, bl = bl0 + adv;
        if (elect_one()) {
          tc_mma_tf32_ts(d_tmem, tmem + kb * 8, bh, idesc, (uint32_t)(kb != 0));
          if (split & 1) {
            tc_mma_tf32_ts(d_tmem, tmem + K + kb * 8, bh, idesc, 1u);
            tc_mma_tf32_ts(d_tmem, tmem + kb * 8, bl, idesc, 1u);
          }
        }
        __syncwarp();
      }
    }
    if (elect_one()) tc_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0u);
    const long long t1 = clock64();
    if (cycles && lane == 0) *cycles = (t1 - t0) / (reps > 0 ? reps : 1);
  }
  __syncthreads();
  tc_fence_after();
  {
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 384;
    uint32_t v[32];
    for (int c = 0; c < 2; ++c) {
      tc_ld32(taddr + c * 32, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) D[(long long)tid * TP_N + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  if (split & 16) {
    // debug: raw registers of 16x256b.x4 loads at lane offsets 0 / 16 of the warp's quarter, column offset 3:
    // D is overwritten with [warp 4][half 2][lane 32][reg 16] (4096 floats of the 8192)
    for (int h = 0; h < 2; ++h) {
      uint32_t v[16];
      tc_ld16x256b_x4(tmem + ((uint32_t)(warp * 32 + 16 * h) << 16) + 384 + 3, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) D[((warp * 2 + h) * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
    }
  }
  if (split & 32) {
    // debug: raw registers of 16x32bx2.x32 loads (second half 32 columns further) at lane offsets 0 / 16 of the
    // warp's quarter, column offset (split >> 8) & 15: D is overwritten with [warp 4][half 2][lane 32][reg 32]
    const uint32_t coff = (split >> 8) & 15;
    for (int h = 0; h < 2; ++h) {
      uint32_t v[32];
      tc_ld16x32bx2_x32(tmem + ((uint32_t)(warp * 32 + 16 * h) << 16) + 384 + coff, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) D[((warp * 2 + h) * 32 + lane) * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace ll

using namespace ll;

extern "C" {

int ll_tc_tf32_probe(const float* A, const float* B, float* D, int kblocks, int split, int reps, long long* cycles,
                     ll_stream_t stream) {
  if (kblocks < 1 || kblocks > TP_MAXKB || reps < 1) return fail(LL_EINVAL, "ll_tc_tf32_probe: kblocks in 1..%d, reps >= 1", TP_MAXKB);
  if (!A || !B || !D) return fail(LL_EINVAL, "ll_tc_tf32_probe: null pointer");
  LL_CUDA_OK(cudaFuncSetAttribute(tc_tf32_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM));
  tc_tf32_probe_kernel<<<1, 128, TP_SMEM, as_stream(stream)>>>(A, B, D, kblocks, split, reps, cycles);
  LL_LAUNCH_OK("tc_tf32_probe_kernel");
  return LL_OK;
}

}  // extern "C"

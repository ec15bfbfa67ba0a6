// lift_tc.cu -- one learned lifting step with the two 16->16 5x5 layers on the tcgen05 tensor cores.
//
//   dout = din + sign * (skip + rw * CNN(skip)),  same function as lift_step_body.cuh (reference:
//   graphs/layers/wavelet_forward_v2.py:58-74, wavelet_inverse_v2.py:76-90, P_block_v2.py:40-55).
//
// 94 % of the step's arithmetic is conv2 / conv3 (16 -> 16 channels, 5x5).  With only 16 output
// channels a pixel-major implicit GEMM would be N = 16; instead the WEIGHTS are the M operand:
//     D[(dx, co)][x] = sum_{dy, ci} W[co][ci][dy][dx] * in[ci][y + dy - 2][x]          (M = 80 of 128, N = 64 columns)
//     out[co][y][x]  = sum_dx D[(dx, co)][x + dx - 2]
// The vertical taps are whole-row shifts of the B operand (a different shared-memory row address per
// MMA: the MMA's own implicit im2col); the horizontal taps become 5 partial planes that the epilogue
// sums with shifted reads.  A (weights, hi and lo halves of both layers: 320 columns) stays resident
// in tensor memory for the whole persistent CTA; B (activations) lives in shared-memory ring buffers
// in the MN-major SWIZZLE_128B_BASE32B layout (atom = 4 channels x 32 columns).
//
// Precision: the path feeds the quantiser, so plain TF32/BF16 is not acceptable.  Every product is
// split 3xTF32 -- A_hi*B_hi + A_lo*B_hi + A_hi*B_lo with hi = rna_tf32(v), lo = rna_tf32(v - hi), FP32
// accumulation in TMEM -- which measures 4e-7 .. 1.4e-6 of the output scale against float64, the same
// as an FP32 FMA chain (tests/test_gpu_tc_probe.py).  30 MMAs (M128 x N64 x K8) per output row and
// layer, 32.5 cycles each when issued straight-line by one elected lane.
//
// Pipeline: the CTA marches down a strip of 52 output columns one row per step.  Four warp roles, each in its own
// step loop (so each keeps only its own loop invariants in its 96 registers), meeting at block barrier 0 once per step:
//   MMA warp 16:          conv2 row t-3 into accumulator (n & 1) -- double-buffered, so its 30 tcgen05.mma start at the
//                         barrier -- then, once the conv3 accumulator of the previous step is drained, conv3 row t-7;
//                         one tcgen05.commit.
//   epilogue warps 0-7:   [E-A] warps 0-2 drain the conv2 accumulator of step n-1, warps 4-6 the conv3 one (quarter 3 is
//                               M padding): a 16-lane TMEM load (tcgen05.ld.16x32bx2) at lane offset 0 / 16 and COLUMN
//                               OFFSET dx hands every thread the dx-shifted samples of its own (co, column) positions,
//                               so the taps dx = 2q, 2q+1 are summed in registers and three pair-sum planes (not five
//                               partial planes) go through shared memory;
//                         [E-B] conv2 row t-4: sum of the pair planes, bias, tanh, hi/lo split -> a2 ring;
//                               conv3 row t-8: sum, bias, + o1 -> a3 ring (fp32, channel pairs interleaved);
//                         skip row t+3 (global loads issued first, stored last).
//   conv1 warps 8-11:     row t (FP32 FFMA2, 4 px x 2 ch per thread, the 25 tap-weight pairs in registers) -> a1 ring
//                         (hi/lo) and o1 ring.
//   conv4 warps 12-15:    conv4 + output row t-11 (FP32 FFMA2, 4 px x 2 ch per thread).
// Tap rows outside the plane read a block of zeros (branch-free); tanh runs two values at a time on the packed FP32
// pipe.  A row produced in step s is consumed in steps > s, ring depths follow.  Every intermediate is forced to 0
// outside the plane (each conv zero-pads its own input).  Timing hooks live in the DBG instantiation only.
#include <cuda_fp16.h>
#include <stdint.h>
#include <string.h>

#include "lift_step_body.cuh"
#include "ll_common.cuh"
#include "tc_ptx.cuh"

namespace ll {

// 8 epilogue warps + 8 SIMT warps + 1 MMA warp.  Warp w issues on SM sub-partition w % 4; the epilogue's accumulator drains
// run on warps 0-2 / 4-6 (TMEM lane quarter = warp % 4), so sub-partition 3 carries ~300 instructions per step less than
// the others.  The MMA warp (~265 per step) is therefore warp 19 -- warps 16-18 exist only to put it there and idle at
// the set-up -- instead of warp 16, which made sub-partition 0 the step's critical path (1800 vs 1535 issue slots).
// The placeholder warps stay resident and take part in every block barrier: letting them exit and counting 544 threads at
// barrier 0 faulted intermittently ("unspecified launch failure", timing dependent).
constexpr int TC_THREADS = 640;
constexpr int TC_MMA_WARP = 19;
#define TC_BAR0() asm volatile("bar.sync 0;" ::: "memory")
constexpr int TC_WO = 52;        // output columns per strip
constexpr int TC_RA = 6;         // a1 / a2 ring rows
constexpr int TC_R3 = 6;         // a3 ring rows
constexpr int TC_RO = 9;         // o1 ring rows
constexpr int TC_RS = 16;        // skip ring rows
constexpr int TC_SLOT = 8192;    // bytes per a1/a2 ring row: [term 2][cig 2][na 2][ka 2][4 ch][128 B]
constexpr int TC_P3 = 60;        // o1 pitch (floats)
constexpr int TC_PA3 = 124;      // a3 channel-pair pitch: 60 columns x 2 channels (+4: pairs land on distinct bank groups)
constexpr int TC_PP = 68;        // partial-plane pitch
constexpr int TC_PS = 72;        // skip pitch
// tensor-memory columns
constexpr int TM_W = 192;        // [layer 2][term 2][dy 5][ci 16] (320 columns, behind the accumulators: the dx-shifted drains
                                 // read up to 4 columns past an accumulator, which must stay inside the allocation)
constexpr int TM_ACC2 = 0, TM_ACC3 = 64, TM_ACC2B = 128;   // conv2 accumulator double-buffered (steps alternate)
// 3xFP16 kernel: weights need 160 columns instead of 320, so the conv3 accumulator is double-buffered as well (no MMA of a
// step waits for a drain) and conv4 (16 -> 1) is a third tensor-core layer, two MMAs per vertical tap.  Its weights live in
// the rows the conv2 regions leave unused -- region 0 (conv2 hi; B = the hi half of a3): rows 96 + dx = w4_hi, rows 101 + dx =
// w4_lo; region 1 (conv2 lo; B = the lo half of a3): rows 106 + dx = w4_hi -- so the three terms of the split land in
// different ROWS of D4 (rows 0..79 of D4 are conv2-weights x a3 garbage nobody reads; conv2's own rows 96.. likewise), all
// fifteen inside the first 16 lanes of TMEM lane quarter 3: ONE 16-lane load drains them (tensor-memory reads run at
// 64 B/clk per SM and the E-A drains already need 40 KB of them per step).
constexpr int TM16_ACC3B = 192, TM16_W = 256, TM16_ACC4 = TM16_W + 160;
static_assert(TM16_ACC4 + 64 <= 512, "tensor-memory budget");
// shared memory (bytes from the 1024-aligned base)
constexpr int TS_RA1 = 0;
constexpr int TS_RA2 = TS_RA1 + TC_RA * TC_SLOT;
constexpr int TS_A3 = TS_RA2 + TC_RA * TC_SLOT;
constexpr int TS_A3_BYTES = TC_R3 * 4096;   // fp32 pair-interleaved ring (TC_R3 * 8 * TC_PA3 * 4 = 23 808 B) or the fp16 operand ring (4 KB rows)
static_assert(TC_R3 * 8 * TC_PA3 * 4 <= TS_A3_BYTES && TS_A3 % 1024 == 0, "a3 ring");
constexpr int TS_O1 = TS_A3 + TS_A3_BYTES;
constexpr int TS_P2 = TS_O1 + TC_RO * 16 * TC_P3 * 4;
constexpr int TS_P3 = TS_P2 + 48 * TC_PP * 4;   // partial planes: [dx pair 3][co 16][TC_PP]
constexpr int TS_SK = TS_P3 + 48 * TC_PP * 4;
constexpr int TS_W = TS_SK + TC_RS * TC_PS * 4;
// small weights (floats): pre[4] W1[tap 25][co 16] b1[16] b2[16] b3[16] W4[tap 25][ci 16] b4[4]
// (channel-pair float2 loads feed packed FFMA2 with a broadcast activation)
constexpr int SW_PRE = 0, SW_W1 = 4, SW_B1 = SW_W1 + 400, SW_B2 = SW_B1 + 16, SW_B3 = SW_B2 + 16, SW_W4 = SW_B3 + 16,
              SW_B4 = SW_W4 + 400, SW_ZERO = SW_B4 + 4, SW_TOTAL = SW_ZERO + 16;   // SW_ZERO: 64 B of zeros = "row outside the plane"
constexpr int TS_P4 = TS_W + SW_TOTAL * 4;          // conv4 partial rows D4[row 15][TC_PP] (3xFP16 kernel)
constexpr int TS_BAR = TS_P4 + 16 * TC_PP * 4;
constexpr int TC_SMEM_BYTES = 1024 + TS_BAR + 64;   // barriers at +0, +16, +32, +40; tensor-memory slot at +24
static_assert(TC_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(TS_RA2 % 1024 == 0 && TS_A3 % 16 == 0 && TS_O1 % 16 == 0 && TS_P2 % 16 == 0 && TS_SK % 16 == 0 && TS_W % 16 == 0 && TS_P4 % 16 == 0 && TS_BAR % 8 == 0, "align");

// byte offset of 4 consecutive columns i..i+3 (i % 4 == 0) of channel ci inside a ring row (hi half)
__device__ __forceinline__ int ring_off(int ci, int i) {
  const int r = ci & 3;
  return (ci >> 3) * 2048 + (i >> 5) * 1024 + ((ci >> 2) & 1) * 512 + r * 128 + ((((i & 31) >> 3) ^ r) << 5) + ((i & 7) << 2);
}

// 128-bit shared load the compiler may not narrow (scalar loads at a 16-byte lane stride are 4-way bank conflicted)
__device__ __forceinline__ float4 lds128(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  asm volatile("" ::"f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));   // all four lanes stay live: ptxas keeps the 128-bit access
  return v;
}

// tanh of two values on the packed FP32 pipe (FMUL2 / FFMA2 / FADD2): |x| < 0.625 -> x + x^3 P(x^2) with the Cephes
// tanhf coefficients (< 2 ulp), else 1 - 2 / (exp(2|x|) + 1) on ex2.approx / rcp.approx (absolute error < 1e-7; the
// scalar tanhf of the CUDA math library has the same structure at twice the instruction count).
__device__ __forceinline__ float2 tanh2(float2 x) {
  const float2 z = __fmul2_rn(x, x);
  float2 q = __ffma2_rn(z, make_float2(-5.70498872745e-3f, -5.70498872745e-3f), make_float2(2.06390887954e-2f, 2.06390887954e-2f));
  q = __ffma2_rn(q, z, make_float2(-5.37397155531e-2f, -5.37397155531e-2f));
  q = __ffma2_rn(q, z, make_float2(1.33314422036e-1f, 1.33314422036e-1f));
  q = __ffma2_rn(q, z, make_float2(-3.33332819422e-1f, -3.33332819422e-1f));
  q = __ffma2_rn(q, __fmul2_rn(z, x), x);
  const float ax = fabsf(x.x), ay = fabsf(x.y);
  float ex, ey;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(ax * 2.885390081777927f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(ay * 2.885390081777927f));
  const float2 e1 = __fadd2_rn(make_float2(ex, ey), make_float2(1.f, 1.f));
  float rx, ry;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(e1.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(e1.y));
  const float2 b = __ffma2_rn(make_float2(rx, ry), make_float2(-2.f, -2.f), make_float2(1.f, 1.f));
  return make_float2(ax < 0.625f ? q.x : copysignf(b.x, x.x), ay < 0.625f ? q.y : copysignf(b.y, x.y));
}

// ---- 3xFP16 variant (LL_LIFT_TC16) -------------------------------------------------------------------------------------
// kind::f16 runs at twice the kind::tf32 rate and an fp16 significand is as wide as a tf32 one (11 bits), so
//   a b  ~=  a_hi b_hi + a_lo b_hi + a_hi b_lo,   hi = fp16(v), lo = fp16(v - hi)
// carries the same 22 bits per operand as the 3xTF32 split at half the tensor time and half the operand bytes.  What fp16
// lacks is exponent range: the lo parts (2^-12 of the value) would be subnormal.  Both operands are therefore pre-scaled by
// a power of two -- activations (tanh outputs, |a| < 1) and weights by 2^8 -- which keeps hi below 65504 and lo normal for
// every |a| >= 1e-3 (smaller values lose bits only below 3e-8 / 2^8 absolute); all three products then land in ONE fp32
// accumulator that holds 2^16 times the sum, and the epilogue multiplies by 2^-16 (exact).  Only used with the tanh
// nonlinearity (linear = 0): without it the activations are unbounded and the launcher keeps the TF32 kernel.
constexpr float TC16_SA = 256.f, TC16_SW = 256.f, TC16_UNSCALE = 1.f / (256.f * 256.f);

// byte offset of 4 consecutive columns i..i+3 (i % 4 == 0) of channel ci inside an fp16 ring row (hi half): MN-major
// SWIZZLE_128B, atom = 8 channels x 64 columns (128 B rows, 16-byte chunks XOR channel), two atoms along K
__device__ __forceinline__ int ring_off16(int ci, int i) {
  const int r = ci & 7;
  return (ci >> 3) * 1024 + r * 128 + ((((i >> 3) ^ r)) << 4) + ((i & 7) << 1);
}

// (SCALE = 1 for the a3 ring: the input of conv4 is not bounded by a tanh, so it keeps the whole fp16 range -- |a3| up to
// 65504 -- and its lo part goes subnormal below |a3| = 0.125, an absolute error of at most 3e-8 per element)
template <int SCALE = 256>
__device__ __forceinline__ void split_store16(uint8_t* row, int ci, int i, const float (&v)[4]) {
  __half hi[4], lo[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float s = SCALE == 1 ? v[k] : v[k] * (float)SCALE;
    hi[k] = __float2half_rn(s);
    lo[k] = __float2half_rn(s - __half2float(hi[k]));
  }
  uint8_t* q = row + ring_off16(ci, i);
  const __half2 h01 = __halves2half2(hi[0], hi[1]), h23 = __halves2half2(hi[2], hi[3]);
  const __half2 l01 = __halves2half2(lo[0], lo[1]), l23 = __halves2half2(lo[2], lo[3]);
  *reinterpret_cast<uint2*>(q) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
  *reinterpret_cast<uint2*>(q + 2048) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

// D[tmem] (+)= A[tmem, fp16 pairs per column] * B[smem desc]; FP16 inputs, FP32 accumulation.
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void split_store(uint8_t* row, int ci, int i, const float (&v)[4]) {
  float hi[4], lo[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    hi[k] = tf32_rna(v[k]);
    lo[k] = tf32_rna(v[k] - hi[k]);
  }
  uint8_t* q = row + ring_off(ci, i);
  *reinterpret_cast<float4*>(q) = make_float4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<float4*>(q + 4096) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

// Timing hooks (scripts/gpu_lift_tc_time.py, gpu_lift_tc_timeline.py) only exist in the DBG instantiation of the
// kernel, which the launcher picks when a debug mode or stamp buffer is set; the product kernel carries none of it.
#define TC_STAMP(k)                                                                      \
  do {                                                                                   \
    if (DBG && p.dbg_buf && blockIdx.x == 0 && n == 40 && lane == 0) p.dbg_buf[warp * 8 + (k)] = clock64(); \
  } while (0)
#define TC_OFF(bit) (DBG && (p.dbg & (bit)))

struct TcSeg {
  int j, b, x0, ya, yb, ny, nx;
};

template <bool DBG, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1) lift_step_tc_kernel(const __grid_constant__ LiftParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  float* A3 = reinterpret_cast<float*>(gen + TS_A3);
  float* O1 = reinterpret_cast<float*>(gen + TS_O1);
  float* P2 = reinterpret_cast<float*>(gen + TS_P2);
  float* P3 = reinterpret_cast<float*>(gen + TS_P3);
  float* SK = reinterpret_cast<float*>(gen + TS_SK);
  float* SW = reinterpret_cast<float*>(gen + TS_W);
  const uint32_t bar_mma = base + TS_BAR, bar_free3 = base + TS_BAR + 16, bar_free4 = base + TS_BAR + 32;
  // bar_seen: every warp that waits for the commit of step n-1 (parity wait) arrives here once it has seen it, and the MMA warp
  // waits for all of them before it commits step n.  Without it a commit that completes early -- nothing but the accumulator
  // drains holds the MMA warp back -- flips the barrier's phase a second time before a slow waiter has looked, the waiter then
  // takes the NEXT phase for the one it is waiting for and the CTA deadlocks (seen as the bounded wait's trap).
  const uint32_t bar_seen = base + TS_BAR + 40;
  constexpr int NSEEN = F16 ? 9 : 8;       // epilogue warps 0-7 (+ the conv4 drain warp 15)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + TS_BAR + 24);
  float* P4 = reinterpret_cast<float*>(gen + TS_P4);
  constexpr int TEND = F16 ? 12 : 11;   // steps past the last output row (the 3xFP16 kernel drains conv4 one step after its MMAs)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup -------------------------------------------------------------------------
  if (tid == 0) {
    mbar_init(bar_mma, 1);
    mbar_init(bar_free3, 4);
    mbar_init(bar_free4, 1);
    mbar_init(bar_seen, NSEEN);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_MMA_WARP) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  for (int i = tid; i < SW_TOTAL; i += TC_THREADS) {
    float v = 0.f;
    if (i < 4) v = p.blob[BL_PRE + i];
    else if (i < SW_B1) v = p.blob[BL_W1 + (i - SW_W1)];                                  // [tap][co]
    else if (i < SW_B2) v = p.blob[BL_B1 + (i - SW_B1)];
    else if (i < SW_B3) v = p.blob[BL_B2 + (i - SW_B2)];
    else if (i < SW_W4) v = p.blob[BL_B3 + (i - SW_B3)];
    else if (i < SW_B4) v = p.blob[BL_W4 + ((i - SW_W4) & 15) * 25 + ((i - SW_W4) >> 4)];   // blob [ci][tap] -> [tap][ci]
    else if (i == SW_B4) v = p.blob[BL_B4];
    SW[i] = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (F16 && warp < 4) {   // fp16 weights: row m = dx*16 + co, 4 regions [layer][term] of 40 columns = [dy 5][ci pair 8]
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t* words = reinterpret_cast<const uint32_t*>(p.blob + BL_TC16);
    const uint32_t* words4 = reinterpret_cast<const uint32_t*>(p.blob + BL_TC16_4);   // conv4: [region 2][row 16][40] -> rows 96..111
    for (int reg = 0; reg < 4; ++reg) {
      const uint32_t* src = tid < 80 ? words + ((long long)reg * 80 + tid) * 40
                                     : (tid >= 96 && tid < 112 && reg < 2) ? words4 + (reg * 16 + (tid - 96)) * 40 : nullptr;
      for (int c = 0; c < 40; c += 8) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = src ? src[c + k] : 0u;
        tmem_st8(trow + TM16_W + reg * 40 + c, v);
      }
    }
    tmem_wait_st();
  }
  if (!F16 && warp < 4) {   // weights -> tensor memory: row m = dx*16 + co (rows 80..127 zero), 4 regions of 80 columns
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int reg = 0; reg < 4; ++reg) {
      const float* src = p.blob + BL_TC + ((long long)reg * 80 + tid) * 80;
      for (int c = 0; c < 80; c += 8) {
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = tid < 80 ? __float_as_uint(src[c + k]) : 0u;
        tmem_st8(trow + TM_W + reg * 80 + c, v);
      }
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // instruction descriptor: D fp32, A/B tf32, A K-major (TMEM), B MN-major, N = 64, M = 128
  // (F16: D fp32, A/B fp16 = format 0, B MN-major plain SWIZZLE_128B, 8-channel atoms 1024 B apart along K)
  const uint32_t idesc = F16 ? (1u << 4) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24)
                             : (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t desc_ra1 = F16 ? umma_desc_mn_sw128(base + TS_RA1, 1024, 1024) : umma_desc_mn_sw128_32b(base + TS_RA1, 1024, 512);
  const uint64_t desc_ra2 = F16 ? umma_desc_mn_sw128(base + TS_RA2, 1024, 1024) : umma_desc_mn_sw128_32b(base + TS_RA2, 1024, 512);
  const uint64_t desc_a3 = umma_desc_mn_sw128(base + TS_A3, 1024, 1024);   // 3xFP16 kernel: a3 ring rows of 4 KB

  const int ncta = gridDim.x;
  const long long lo_u = p.total_units * (long long)blockIdx.x / ncta;
  const long long hi_u = p.total_units * (long long)(blockIdx.x + 1) / ncta;
  uint32_t n = 0;   // global step counter (mbarrier phases)

  for (long long u = lo_u; u < hi_u;) {
    // ---- segment: a run of consecutive 8-row chunks of one strip -------------------------------
    TcSeg s;
    {
      long long v = u;
      int j = 0;
      if (v >= p.units[0]) { v -= p.units[0]; j = 1; }
      const int nch = p.nchunks[j], nst = p.nstrips[j];
      const int c = (int)(v % nch);
      const long long w = v / nch;
      s.j = j;
      s.b = (int)(w / nst);
      s.x0 = (int)(w % nst) * TC_WO;
      s.ny = p.job[j].ny;
      s.nx = p.job[j].nx;
      long long cnt = nch - c;
      if (cnt > hi_u - u) cnt = hi_u - u;
      s.ya = c * LS_R;
      s.yb = s.ya + (int)cnt * LS_R;
      if (s.yb > s.ny) s.yb = s.ny;
      u += cnt;
    }
    const ll_lift_job& J = p.job[s.j];
    const float* srcb = J.src.ptr + (long long)s.b * J.src.sb;
    const int a1_lo = max(0, s.ya - 6), a1_hi = min(s.ny, s.yb + 6);
    const int a2_lo = max(0, s.ya - 4), a2_hi = min(s.ny, s.yb + 4);
    const int a3_lo = max(0, s.ya - 2), a3_hi = min(s.ny, s.yb + 2);
    const int sk_lo = max(0, s.ya - 8), sk_hi = min(s.ny, s.yb + 8);

    auto skip_row = [&](int rr, int j) {   // skip ring row rr, column j (plane column x0 - 8 + j)
      const int c = s.x0 - 8 + j;
      float v = 0.f;
      if (c >= 0 && c < s.nx) {
        const float* q = srcb + (long long)rr * J.src.sy + (long long)c * J.src.sx;
        const float s0 = rr > 0 ? q[-J.src.sy] : 0.f;
        const float s1 = q[0];
        const float s2 = rr + 1 < s.ny ? q[J.src.sy] : 0.f;
        v = fmaf(SW[SW_PRE + 2], s2, fmaf(SW[SW_PRE + 1], s1, __fmul_rn(SW[SW_PRE + 0], s0)));
      }
      SK[(rr & (TC_RS - 1)) * TC_PS + j] = v;
    };

    // conv4 + output, 4 pixels x 2 input channels per thread (c4 in [0,104): ci pair = c4 & 7, pixel quad = c4 >> 3).
    // a3 is stored with the two channels of a pair interleaved ([row][pair][col][2]) so (a3[2cp], a3[2cp+1]) of a
    // column is one aligned register pair: packed FFMA2 against the weight pair, no operand shuffling.
    auto conv4_out = [&](int r4, int c4, bool active, const float (&dv)[4]) {
      const int cp = c4 & 7, pq = c4 >> 3;
      float2 acc[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = make_float2(0.f, 0.f);
      if (active) {
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) {
          // rows outside the plane read a block of zeros instead of being skipped: no branch, so all the row and
          // weight loads of the 5 taps can be in flight together (the loop used to pay 5 shared-memory round trips)
          const int rr = r4 + dy - 2;
          const float* ar = (rr >= 0 && rr < s.ny) ? A3 + ((rr % TC_R3) * 8 + cp) * TC_PA3 + 8 * pq : SW + SW_ZERO;   // columns 4pq .. 4pq+7, 2 channels each
          const float4 v0 = *reinterpret_cast<const float4*>(ar), v1 = *reinterpret_cast<const float4*>(ar + 4),
                       v2 = *reinterpret_cast<const float4*>(ar + 8), v3 = *reinterpret_cast<const float4*>(ar + 12);
          const float2 col[8] = {make_float2(v0.x, v0.y), make_float2(v0.z, v0.w), make_float2(v1.x, v1.y), make_float2(v1.z, v1.w),
                                 make_float2(v2.x, v2.y), make_float2(v2.z, v2.w), make_float2(v3.x, v3.y), make_float2(v3.z, v3.w)};
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {    // (the 25 weight pairs do not fit this role's registers next to the 5x16 window)
            const float2 w = *reinterpret_cast<const float2*>(SW + SW_W4 + (dy * 5 + dx) * 16 + 2 * cp);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] = __ffma2_rn(col[k + dx], w, acc[k]);
          }
        }
      }
      float a4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a4[k] = acc[k].x + acc[k].y;
        a4[k] += __shfl_xor_sync(0xffffffffu, a4[k], 1);
        a4[k] += __shfl_xor_sync(0xffffffffu, a4[k], 2);
        a4[k] += __shfl_xor_sync(0xffffffffu, a4[k], 4);
      }
      if (active && cp == 0) {
        const float b4 = SW[SW_B4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int jx = 4 * pq + k, c = s.x0 + jx;
          if (c < s.nx) {
            const float net = __fadd_rn(a4[k], b4);
            const float sk = SK[(r4 & (TC_RS - 1)) * TC_PS + jx + 8];
            const float tn = __fmul_rn(net, p.rw);
            const float o = p.sign > 0.f ? __fadd_rn(__fadd_rn(dv[k], sk), tn)
                                         : p.sign < 0.f ? __fadd_rn(__fadd_rn(dv[k], -sk), -tn) : net;
            J.dout.ptr[(long long)s.b * J.dout.sb + (long long)r4 * J.dout.sy + (long long)c * J.dout.sx] = o;
          }
        }
      }
    };
    // Branch-free: lanes that have nothing to load read the plane's first sample instead (a divergent region around
    // the loads makes the warp wait for them at its reconvergence point, ~650 cycles at the start of every step).
    const float* din_b = J.din.ptr + (long long)s.b * J.din.sb;
    auto load_din = [&](int r4, int c4, bool active, float (&dv)[4]) {
      const float* row = din_b + (long long)r4 * J.din.sy;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = s.x0 + 4 * (c4 >> 3) + k;
        const bool ok = active && (c4 & 7) == 0 && c < s.nx;
        dv[k] = *(ok ? row + (long long)c * J.din.sx : din_b);
      }
    };

    // prologue: skip rows ya-8 .. ya-4 (row ya-3 is produced by the first step)
    if (warp < 16) {
      for (int e = tid; e < 5 * 68; e += 512) {
        const int rr = s.ya - 8 + e / 68;
        if (rr >= sk_lo && rr < sk_hi) skip_row(rr, e % 68);
      }
    }
    TC_BAR0();

    // One step loop per warp role (instead of one loop with a role switch inside): each role keeps only its own
    // loop invariants in registers.  All three loops run the same steps and meet at barrier 0 once per step.
    if (warp >= 16 && warp != TC_MMA_WARP) {          // placeholder warps (see TC_MMA_WARP): only the step barriers
      for (int t = s.ya - 6; t < s.yb + TEND; ++t, ++n) TC_BAR0();
    } else if (warp == TC_MMA_WARP) {
      for (int t = s.ya - 6; t < s.yb + TEND; ++t, ++n) {
        const int r2m = t - 3, r3m = t - 7;      // rows whose MMAs are issued in this step
        const int r2e = t - 4, r3e = t - 8;      // rows whose accumulators are drained in this step
        const bool e2 = r2e >= a2_lo && r2e < a2_hi, e3 = r3e >= a3_lo && r3e < a3_hi;
          // ======================= MMA warp =======================
          // conv2 alternates between two accumulators: the one of step n was drained in step n-1 (behind the
          // end-of-step barrier), so its MMAs start at once; only conv3 waits for this step's drain.
          TC_STAMP(0);
          if (DBG && p.dbg_buf && blockIdx.x == 0 && lane == 0 && (n == 40 || n == 240)) {   // wall clock beside the cycle counter: SM clock under load
            long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            p.dbg_buf[warp * 8 + (n == 40 ? 3 : 4)] = gt;
            if (n == 240) p.dbg_buf[warp * 8 + 5] = clock64();
          }
          tc_fence_after();
          const uint32_t acc2 = tmem + TM_ACC2 + (n & 1) * (TM_ACC2B - TM_ACC2);
          const bool m2 = r2m >= a2_lo && r2m < a2_hi, m3 = r3m >= a3_lo && r3m < a3_hi;
          if (F16) {
            // Both accumulators of step n were drained in step n-1: all 30 MMAs of conv2 / conv3 go out at the barrier, the
            // two layers' (independent) accumulation chains interleaved tap by tap.
            const uint32_t acc3 = tmem + ((n & 1) ? TM16_ACC3B : TM_ACC3);
            if (elect_one()) {
              if (!TC_OFF(1)) {
                uint32_t c2 = 0, c3 = 0;
#pragma unroll
                for (int dy = 0; dy < 5; ++dy) {
                  const int rr2 = r2m + dy - 2, rr3 = r3m + dy - 2;
                  const bool ok2 = m2 && rr2 >= 0 && rr2 < s.ny, ok3 = m3 && rr3 >= 0 && rr3 < s.ny;
                  const uint64_t bd2 = desc_ra1 + (uint32_t)(((rr2 + TC_RA) % TC_RA) * (TC_SLOT >> 4));
                  const uint64_t bd3 = desc_ra2 + (uint32_t)(((rr3 + TC_RA) % TC_RA) * (TC_SLOT >> 4));
                  const uint32_t ah2 = tmem + TM16_W + 0 * 80 + dy * 8, al2 = ah2 + 40;
                  const uint32_t ah3 = tmem + TM16_W + 1 * 80 + dy * 8, al3 = ah3 + 40;
                  if (ok2) tc_mma_f16_ts(acc2, ah2, bd2, idesc, c2);                   // hi * hi
                  if (ok3) tc_mma_f16_ts(acc3, ah3, bd3, idesc, c3);
                  if (ok2) tc_mma_f16_ts(acc2, al2, bd2, idesc, 1u);                   // lo * hi
                  if (ok3) tc_mma_f16_ts(acc3, al3, bd3, idesc, 1u);
                  if (ok2) tc_mma_f16_ts(acc2, ah2, bd2 + (2048 >> 4), idesc, 1u);     // hi * lo
                  if (ok3) tc_mma_f16_ts(acc3, ah3, bd3 + (2048 >> 4), idesc, 1u);
                  if (ok2) c2 = 1;
                  if (ok3) c3 = 1;
                }
              }
            }
            __syncwarp();
            TC_STAMP(1);
            // conv4 row t-11 (its a3 rows were completed in step t-1) once the previous step's D4 has been drained
            const int r4m = t - 11;
            const bool m4 = r4m >= s.ya && r4m < s.yb;
            mbar_wait(bar_free4, n & 1);
            tc_fence_after();
            if (elect_one()) {
              if (m4 && !TC_OFF(1)) {
                uint32_t acc = 0;
#pragma unroll
                for (int dy = 0; dy < 5; ++dy) {
                  const int rr = r4m + dy - 2;
                  if (rr >= 0 && rr < s.ny) {
                    const uint64_t bd = desc_a3 + (uint32_t)((rr % TC_R3) * (4096 >> 4));
                    const uint32_t a0 = tmem + TM16_W + dy * 8, a1 = a0 + 40;
                    tc_mma_f16_ts(tmem + TM16_ACC4, a0, bd, idesc, acc);                  // rows 96 + dx: hi * hi, rows 101 + dx: lo * hi
                    tc_mma_f16_ts(tmem + TM16_ACC4, a1, bd + (2048 >> 4), idesc, 1u);     // rows 106 + dx: hi * lo
                    acc = 1;
                  }
                }
              }
            }
            __syncwarp();
            mbar_wait(bar_seen, n & 1);
            if (elect_one()) tc_commit(bar_mma);
          } else {
          if (elect_one()) {
            if (m2 && !TC_OFF(1)) {
              uint32_t acc = 0;
#pragma unroll
              for (int dy = 0; dy < 5; ++dy) {
                const int rr = r2m + dy - 2;
                if (rr >= 0 && rr < s.ny) {
                  const uint64_t bd = desc_ra1 + (uint32_t)((rr % TC_RA) * (TC_SLOT >> 4));
                  const uint32_t ah = tmem + TM_W + 0 * 160 + dy * 16, al = ah + 80;
                  tc_mma_tf32_ts(acc2, ah, bd, idesc, acc);                       // hi * hi, ci 0-7
                  tc_mma_tf32_ts(acc2, al, bd, idesc, 1u);                        // lo * hi
                  tc_mma_tf32_ts(acc2, ah, bd + (4096 >> 4), idesc, 1u);          // hi * lo
                  tc_mma_tf32_ts(acc2, ah + 8, bd + (2048 >> 4), idesc, 1u);      // ci 8-15
                  tc_mma_tf32_ts(acc2, al + 8, bd + (2048 >> 4), idesc, 1u);
                  tc_mma_tf32_ts(acc2, ah + 8, bd + ((4096 + 2048) >> 4), idesc, 1u);
                  acc = 1;
                }
              }
            }
          }
          TC_STAMP(1);
          mbar_wait(bar_free3, n & 1);       // only conv3 waits for this step's drain (one conv3 accumulator)
          tc_fence_after();
          if (elect_one()) {
            if (m3 && !TC_OFF(1)) {
              uint32_t acc = 0;
#pragma unroll
              for (int dy = 0; dy < 5; ++dy) {
                const int rr = r3m + dy - 2;
                if (rr >= 0 && rr < s.ny) {
                  const uint64_t bd = desc_ra2 + (uint32_t)((rr % TC_RA) * (TC_SLOT >> 4));
                  const uint32_t ah = tmem + TM_W + 1 * 160 + dy * 16, al = ah + 80;
                  tc_mma_tf32_ts(tmem + TM_ACC3, ah, bd, idesc, acc);
                  tc_mma_tf32_ts(tmem + TM_ACC3, al, bd, idesc, 1u);
                  tc_mma_tf32_ts(tmem + TM_ACC3, ah, bd + (4096 >> 4), idesc, 1u);
                  tc_mma_tf32_ts(tmem + TM_ACC3, ah + 8, bd + (2048 >> 4), idesc, 1u);
                  tc_mma_tf32_ts(tmem + TM_ACC3, al + 8, bd + (2048 >> 4), idesc, 1u);
                  tc_mma_tf32_ts(tmem + TM_ACC3, ah + 8, bd + ((4096 + 2048) >> 4), idesc, 1u);
                  acc = 1;
                }
              }
            }
          }
          __syncwarp();
          mbar_wait(bar_seen, n & 1);
          if (elect_one()) tc_commit(bar_mma);
          }
          __syncwarp();
          TC_STAMP(2);
        TC_STAMP(6);
        TC_BAR0();   // end of step: every role loop issues exactly one per step
        TC_STAMP(7);
      }
    } else if (warp < 8) {
      for (int t = s.ya - 6; t < s.yb + TEND; ++t, ++n) {
        const int r2m = t - 3, r3m = t - 7;      // rows whose MMAs are issued in this step
        const int r2e = t - 4, r3e = t - 8;      // rows whose accumulators are drained in this step
        const bool e2 = r2e >= a2_lo && r2e < a2_hi, e3 = r3e >= a3_lo && r3e < a3_hi;
          // ======================= epilogue warps =======================
          // skip row t+3 (threads 128..195): global loads first, shared-memory store at the end of the step
          const int rs = t + 3;
          const bool do_sk = tid >= 128 && tid < 196 && rs >= sk_lo && rs < sk_hi;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f;
          bool sk_in = false;
          if (do_sk) {
            const int c = s.x0 - 8 + (tid - 128);
            sk_in = c >= 0 && c < s.nx;
            if (sk_in) {
              const float* q = srcb + (long long)rs * J.src.sy + (long long)c * J.src.sx;
              s0 = rs > 0 ? q[-J.src.sy] : 0.f;
              s1 = q[0];
              s2 = rs + 1 < s.ny ? q[J.src.sy] : 0.f;
            }
          }
          TC_STAMP(0);
          // ---- E-A: accumulators of the previous step -> partial planes ----
          if (n > 0) mbar_wait(bar_mma, (n - 1) & 1);
          if (lane == 0) mbar_arrive(bar_seen);
          tc_fence_after();
          TC_STAMP(1);
          // Warps 0-3 drain the conv2 accumulator, warps 4-7 the conv3 one; warp quarter q owns accumulator lanes
          // 32q .. 32q+31 = the rows of dx = 2q (lanes 0-15) and dx = 2q+1 (lanes 16-31).  A 16-lane load
          // (16x32bx2: threads 0-15 <- columns c .. c+31 of lane t, threads 16-31 <- columns c+32 .. c+63 of lane t-16)
          // at column offset dx hands every thread the dx-shifted samples of ITS (co, column) positions, so the two
          // taps of a quarter are summed in registers and only 3 pair-sum planes go through shared memory.
          {
            const int q = warp & 3, l3 = warp >> 2;
            const bool on = (l3 ? e3 : e2) && q < 3 && !TC_OFF(16);
            if (on) {
              const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + (l3 ? (F16 ? ((n & 1) ? TM_ACC3 : TM16_ACC3B) : TM_ACC3) : ((n & 1) ? TM_ACC2 : TM_ACC2B)) + 2 * q;   // buffers of step n-1
              float* o = (l3 ? P3 : P2) + (q * 16 + (lane & 15)) * TC_PP + 32 * (lane >> 4);
#pragma unroll
              for (int r = 0; r < 4; ++r) {            // four rounds of 8 columns per half keep 16 registers live
                uint32_t va[8], vb[8];
                tc_ld16x32bx2_x8(ta + 8 * r, va);
                if (q < 2) tc_ld16x32bx2_x8(ta + (16u << 16) + 1 + 8 * r, vb);
                tc_wait_ld();
#pragma unroll
                for (int k = 0; k < 8; k += 4) {
                  float4 f = make_float4(__uint_as_float(va[k]), __uint_as_float(va[k + 1]), __uint_as_float(va[k + 2]), __uint_as_float(va[k + 3]));
                  if (q < 2) {
                    f.x = __fadd_rn(f.x, __uint_as_float(vb[k]));
                    f.y = __fadd_rn(f.y, __uint_as_float(vb[k + 1]));
                    f.z = __fadd_rn(f.z, __uint_as_float(vb[k + 2]));
                    f.w = __fadd_rn(f.w, __uint_as_float(vb[k + 3]));
                  }
                  *reinterpret_cast<float4*>(o + 8 * r + k) = f;
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (l3 && lane == 0) mbar_arrive(bar_free3);   // this warp's share of the conv3 accumulator is drained
            if (F16 && l3) asm volatile("bar.arrive 3, 256;" ::: "memory");   // conv3 pair planes complete: warps 12-15 take them from here
          }
          TC_STAMP(2);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          TC_STAMP(3);

          // ---- E-B: conv2 row r2e -> a2 ring; conv3 row r3e -> a3 ring ----
          {
            const int co = tid >> 4, xq = tid & 15, i0 = 4 * xq;
            if (e2 && !TC_OFF(2)) {
              const float* pr = P2 + co * TC_PP + i0;     // pair-sum planes already hold the dx-shifted samples
              const float4 s01 = lds128(pr), s23 = lds128(pr + 16 * TC_PP), s4 = lds128(pr + 32 * TC_PP);
              const float us = F16 ? TC16_UNSCALE : 1.f;   // the fp16 accumulators hold 2^16 x the sum (exact to undo)
              const float sacc[4] = {__fadd_rn(__fadd_rn(s01.x, s23.x), s4.x) * us, __fadd_rn(__fadd_rn(s01.y, s23.y), s4.y) * us,
                                     __fadd_rn(__fadd_rn(s01.z, s23.z), s4.z) * us, __fadd_rn(__fadd_rn(s01.w, s23.w), s4.w) * us};
              const float bias = SW[SW_B2 + co];
              float v[4];
              {
                float2 x01 = make_float2(__fadd_rn(sacc[0], bias), __fadd_rn(sacc[1], bias));
                float2 x23 = make_float2(__fadd_rn(sacc[2], bias), __fadd_rn(sacc[3], bias));
                if (!p.linear) {
                  x01 = tanh2(x01);
                  x23 = tanh2(x23);
                }
                const float x[4] = {x01.x, x01.y, x23.x, x23.y};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int c = s.x0 - 4 + i0 + k;
                  v[k] = (c >= 0 && c < s.nx && i0 + k < 60) ? x[k] : 0.f;
                }
              }
              if (F16) split_store16(gen + TS_RA2 + (r2e % TC_RA) * TC_SLOT, co, i0, v);
              else split_store(gen + TS_RA2 + (r2e % TC_RA) * TC_SLOT, co, i0, v);
            }
            if (!F16 && e3 && !TC_OFF(2)) {       // (xq = 14, 15 compute on stale columns and store nothing)
              const float* pr = P3 + co * TC_PP + i0;     // pair-sum planes already hold the dx-shifted samples
              const float4 s01 = lds128(pr), s23 = lds128(pr + 16 * TC_PP), s4 = lds128(pr + 32 * TC_PP);
              const float us = F16 ? TC16_UNSCALE : 1.f;
              const float sacc[4] = {__fadd_rn(__fadd_rn(s01.x, s23.x), s4.x) * us, __fadd_rn(__fadd_rn(s01.y, s23.y), s4.y) * us,
                                     __fadd_rn(__fadd_rn(s01.z, s23.z), s4.z) * us, __fadd_rn(__fadd_rn(s01.w, s23.w), s4.w) * us};
              const float bias = SW[SW_B3 + co];
              const float4 o1 = *reinterpret_cast<const float4*>(O1 + ((r3e % TC_RO) * 16 + co) * TC_P3 + i0);
              const float o1v[4] = {o1.x, o1.y, o1.z, o1.w};
              float v[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int c = s.x0 - 2 + i0 + k;
                const float x = __fadd_rn(__fadd_rn(sacc[k], bias), o1v[k]);
                v[k] = (c >= 0 && c < s.nx) ? x : 0.f;
              }
              {
              // the partner lane (xor 16) holds the other channel of the pair for the same 4 columns
              float o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) o[k] = __shfl_xor_sync(0xffffffffu, v[k], 16);
              float* dst = A3 + ((r3e % TC_R3) * 8 + (co >> 1)) * TC_PA3 + 2 * i0;
              if (xq < 14) {
                if (co & 1) *reinterpret_cast<float4*>(dst + 4) = make_float4(o[2], v[2], o[3], v[3]);   // columns i0+2, i0+3
                else *reinterpret_cast<float4*>(dst) = make_float4(v[0], o[0], v[1], o[1]);               // columns i0, i0+1
              }
              }
            }
          }
          TC_STAMP(4);
          if (do_sk) {
            float v = 0.f;
            if (sk_in) v = fmaf(SW[SW_PRE + 2], s2, fmaf(SW[SW_PRE + 1], s1, __fmul_rn(SW[SW_PRE + 0], s0)));
            SK[(rs & (TC_RS - 1)) * TC_PS + (tid - 128)] = v;
          }
          TC_STAMP(5);
          fence_proxy_async();   // a2 (and a3) ring writes -> visible to the tensor core's operand reads
        TC_STAMP(6);
        TC_BAR0();   // end of step: every role loop issues exactly one per step
        TC_STAMP(7);
      }
    } else if (warp < 12) {
      // ======================= conv1 warps 8-11 =======================
      // conv1 row t -> a1 ring (hi/lo) and o1 ring: thread = 4 pixels x 2 channels (FFMA2); 16 consecutive lanes = the
      // 16 pixel quads of one channel pair.  The 25 tap weights of the pair live in registers for the whole segment.
      const int st = tid - 256;
      const int cp = st >> 4, i1 = 4 * (st & 15);
      float2 wr[25];
#pragma unroll
      for (int k = 0; k < 25; ++k) wr[k] = *reinterpret_cast<const float2*>(SW + SW_W1 + k * 16 + 2 * cp);
      const float2 b1 = *reinterpret_cast<const float2*>(SW + SW_B1 + 2 * cp);
      for (int t = s.ya - 6; t < s.yb + TEND; ++t, ++n) {
        TC_STAMP(0);
        if (t >= a1_lo && t < a1_hi && !TC_OFF(4)) {
          float2 acc[4] = {b1, b1, b1, b1};
#pragma unroll
          for (int dy = 0; dy < 5; ++dy) {
            const int rr = t + dy - 2;            // rows outside the plane: zeros, no branch (see conv4_out)
            const float* sr = (rr >= 0 && rr < s.ny) ? SK + (rr & (TC_RS - 1)) * TC_PS + i1 : SW + SW_ZERO;
            const float4 a = *reinterpret_cast<const float4*>(sr), b = *reinterpret_cast<const float4*>(sr + 4);
            const float win[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) {
#pragma unroll
              for (int k = 0; k < 4; ++k) acc[k] = __ffma2_rn(make_float2(win[k + dx], win[k + dx]), wr[dy * 5 + dx], acc[k]);
            }
          }
          float2 th[4];                       // tanh of the channel pair, two values per packed instruction
#pragma unroll
          for (int k = 0; k < 4; ++k) th[k] = p.linear ? acc[k] : tanh2(acc[k]);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float av[4], ov[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int c = s.x0 - 6 + i1 + k;
              const bool in = c >= 0 && c < s.nx;
              ov[k] = in ? (h ? acc[k].y : acc[k].x) : 0.f;
              av[k] = in ? (h ? th[k].y : th[k].x) : 0.f;
            }
            const int co = 2 * cp + h;
            if (F16) split_store16(gen + TS_RA1 + (t % TC_RA) * TC_SLOT, co, i1, av);
            else split_store(gen + TS_RA1 + (t % TC_RA) * TC_SLOT, co, i1, av);
            if (i1 >= 4 && i1 < 60)
              *reinterpret_cast<float4*>(O1 + ((t % TC_RO) * 16 + co) * TC_P3 + i1 - 4) = make_float4(ov[0], ov[1], ov[2], ov[3]);
          }
        }
        TC_STAMP(1);
        fence_proxy_async();   // a1 ring writes -> visible to the tensor core's operand reads
        TC_STAMP(6);
        TC_BAR0();   // end of step: every role loop issues exactly one per step
        TC_STAMP(7);
      }
    } else {
      // ======================= conv4 + output warps 12-15 (104 active threads) =======================
      const int c4 = tid - 384;
      const bool act4 = c4 < 104;
      if (F16) {
        // 3xFP16 kernel.  (1) conv4 ran on the tensor cores in the previous step: warp 15 (TMEM lane quarter 3) moves the
        // fifteen rows of D4 that carry the three terms to shared memory, thread c4 < 52 then owns output column c4:
        // net = sum_dx (lo*hi + hi*lo + hi*hi)[dx][c4 + dx].  (2) The conv3 half of E-B: pair planes -> + bias + o1 -> a3 operand
        // ring (UNSCALED fp16 hi/lo: the input of conv4 is not bounded by a tanh), two (channel, column quad) items per thread.
        const bool col4 = c4 < TC_WO && s.x0 + c4 < s.nx;
        for (int t = s.ya - 6; t < s.yb + TEND; ++t, ++n) {
          const int r4 = t - 12, r3e = t - 8;
          const bool do4 = r4 >= s.ya && r4 < s.yb && !TC_OFF(8);
          const bool e3 = r3e >= a3_lo && r3e < a3_hi;
          const float dv = *((do4 && col4) ? din_b + (long long)r4 * J.din.sy + (long long)(s.x0 + c4) * J.din.sx : din_b);
          TC_STAMP(0);
          if (warp == 15) {
            if (n > 0) mbar_wait(bar_mma, (n - 1) & 1);
            if (lane == 0) mbar_arrive(bar_seen);
            tc_fence_after();
            if (do4) {
              // threads 0-15 <- D4 row 96 + t, columns 0..31; threads 16-31 <- row 96 + (t - 16), columns 32..63
              uint32_t v[32];
              tc_ld16x32bx2_x32(tmem + ((uint32_t)96 << 16) + TM16_ACC4, v);
              tc_wait_ld();
              if ((lane & 15) < 15) {
                float* o = P4 + (lane & 15) * TC_PP + 32 * (lane >> 4);
#pragma unroll
                for (int k = 0; k < 32; k += 4)
                  *reinterpret_cast<float4*>(o + k) = make_float4(__uint_as_float(v[k]), __uint_as_float(v[k + 1]), __uint_as_float(v[k + 2]), __uint_as_float(v[k + 3]));
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_free4);
          }
          TC_STAMP(1);
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (do4 && col4) {
            float a4 = 0.f;
#pragma unroll
            for (int dx = 0; dx < 5; ++dx)      // small terms first, then the hi * hi row
              a4 = __fadd_rn(a4, __fadd_rn(__fadd_rn(P4[(5 + dx) * TC_PP + c4 + dx], P4[(10 + dx) * TC_PP + c4 + dx]), P4[dx * TC_PP + c4 + dx]));
            const float net = __fadd_rn(a4 * (1.f / TC16_SW), SW[SW_B4]);
            const float sk = SK[(r4 & (TC_RS - 1)) * TC_PS + c4 + 8];
            const float tn = __fmul_rn(net, p.rw);
            const float o = p.sign > 0.f ? __fadd_rn(__fadd_rn(dv, sk), tn)
                                         : p.sign < 0.f ? __fadd_rn(__fadd_rn(dv, -sk), -tn) : net;
            J.dout.ptr[(long long)s.b * J.dout.sb + (long long)r4 * J.dout.sy + (long long)(s.x0 + c4) * J.dout.sx] = o;
          }
          TC_STAMP(2);
          asm volatile("bar.sync 3, 256;" ::: "memory");     // conv3 pair planes of this step (E-A of warps 4-6)
          if (e3 && !TC_OFF(2)) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
              const int item = c4 + 128 * it, co = item >> 4, xq = item & 15, i0 = 4 * xq;
              const float* pr = P3 + co * TC_PP + i0;     // pair-sum planes already hold the dx-shifted samples
              const float4 s01 = lds128(pr), s23 = lds128(pr + 16 * TC_PP), s4 = lds128(pr + 32 * TC_PP);
              const float sacc[4] = {__fadd_rn(__fadd_rn(s01.x, s23.x), s4.x) * TC16_UNSCALE, __fadd_rn(__fadd_rn(s01.y, s23.y), s4.y) * TC16_UNSCALE,
                                     __fadd_rn(__fadd_rn(s01.z, s23.z), s4.z) * TC16_UNSCALE, __fadd_rn(__fadd_rn(s01.w, s23.w), s4.w) * TC16_UNSCALE};
              const float bias = SW[SW_B3 + co];
              const float4 o1 = *reinterpret_cast<const float4*>(O1 + ((r3e % TC_RO) * 16 + co) * TC_P3 + i0);
              const float o1v[4] = {o1.x, o1.y, o1.z, o1.w};
              float v[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int c = s.x0 - 2 + i0 + k;
                const float x = __fadd_rn(__fadd_rn(sacc[k], bias), o1v[k]);
                v[k] = (c >= 0 && c < s.nx && xq < 14) ? x : 0.f;   // columns 56..63 only feed accumulator columns nobody reads
              }
              split_store16<1>(gen + TS_A3 + (r3e % TC_R3) * 4096, co, i0, v);
            }
          }
          TC_STAMP(3);
          fence_proxy_async();   // a3 ring writes -> visible to the tensor core's operand reads
          TC_STAMP(6);
          TC_BAR0();
          TC_STAMP(7);
        }
      } else
      for (int t = s.ya - 6; t < s.yb + TEND; ++t, ++n) {
        const int r4 = t - 11;
        const bool do4 = r4 >= s.ya && r4 < s.yb && !TC_OFF(8);
        float dv[4];
        load_din(r4, c4, do4 && act4, dv);     // global loads first; used at the end of conv4
        TC_STAMP(0);
        if (do4) conv4_out(r4, c4, act4, dv);
        TC_STAMP(2);
        TC_STAMP(6);
        TC_BAR0();   // end of step: every role loop issues exactly one per step
        TC_STAMP(7);
      }
    }
  }

  // drain: the last step's commit must have landed before tensor memory is released
  if (warp < 8 && n > 0) mbar_wait(bar_mma, (n - 1) & 1);
  tc_fence_before();
  TC_BAR0();
  if (warp == TC_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// (region = layer*2 + term, row m = dx*16 + co, col = dy*16 + ci) of the tensor-core weight block
__global__ void pack_lift_tc_kernel(const float* __restrict__ w2, const float* __restrict__ w3, float* __restrict__ blob) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * 80 * 80; i += gridDim.x * blockDim.x) {
    const int col = i % 80, m = (i / 80) % 80, reg = i / 6400;
    const int dy = col / 16, ci = col % 16, dx = m / 16, co = m % 16;
    const float w = ((reg >> 1) ? w3 : w2)[(co * 16 + ci) * 25 + dy * 5 + dx];
    const float hi = tf32_rna(w);
    blob[BL_TC + i] = (reg & 1) ? tf32_rna(w - hi) : hi;
  }
}

// fp16 weight block (LL_LIFT_TC16): [region = layer*2 + term][row m = dx*16 + co][word = dy*8 + ci/2], each word the fp16
// pair (ci even, ci odd) of hi = fp16(256 w) or lo = fp16(256 w - hi)
__global__ void pack_lift_tc16_kernel(const float* __restrict__ w2, const float* __restrict__ w3, float* __restrict__ blob) {
  uint32_t* words = reinterpret_cast<uint32_t*>(blob + BL_TC16);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * 80 * 40; i += gridDim.x * blockDim.x) {
    const int word = i % 40, m = (i / 40) % 80, reg = i / 3200;
    const int dy = word / 8, cp = word % 8, dx = m / 16, co = m % 16;
    __half h[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float w = ((reg >> 1) ? w3 : w2)[(co * 16 + 2 * cp + e) * 25 + dy * 5 + dx] * TC16_SW;
      const __half hi = __float2half_rn(w);
      h[e] = (reg & 1) ? __float2half_rn(w - __half2float(hi)) : hi;
    }
    const __half2 pr = __halves2half2(h[0], h[1]);
    words[i] = *reinterpret_cast<const uint32_t*>(&pr);
  }
}

// conv4 block of the 3xFP16 kernel: [region 2][row 16][word = dy*8 + ci/2], fp16 pairs (ci even, ci odd).  Region 0 multiplies
// the hi half of a3: rows dx = hi = fp16(256 w), rows 5 + dx = lo = fp16(256 w - hi); region 1 multiplies the lo half: rows
// 10 + dx = hi; every other row is zero.  (The kernel puts them at tensor-memory rows 96..111.)  w4 is the torch (1,16,5,5) tensor.
__global__ void pack_lift_tc16_conv4_kernel(const float* __restrict__ w4, float* __restrict__ blob) {
  uint32_t* words = reinterpret_cast<uint32_t*>(blob + BL_TC16_4);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * 16 * 40; i += gridDim.x * blockDim.x) {
    const int word = i % 40, row = (i / 40) % 16, reg = i / 640;
    const int dy = word / 8, cp = word % 8, grp = row / 5, dx = row % 5;
    const bool live = row < 15 && (reg == 0 ? grp < 2 : grp == 2);
    __half h[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float w = live ? w4[(2 * cp + e) * 25 + dy * 5 + dx] * TC16_SW : 0.f;
      const __half hi = __float2half_rn(w);
      h[e] = grp == 1 ? __float2half_rn(w - __half2float(hi)) : hi;
    }
    const __half2 pr = __halves2half2(h[0], h[1]);
    words[i] = *reinterpret_cast<const uint32_t*>(&pr);
  }
}

int launch_lift_step_tc(const LiftParams& p0, cudaStream_t stream) {
  LiftParams p = p0;
  p.total_units = 0;
  for (int j = 0; j < 2; ++j) {
    if (j < p.njobs) {
      p.nstrips[j] = (p.job[j].nx + TC_WO - 1) / TC_WO;
      p.nchunks[j] = (p.job[j].ny + LS_R - 1) / LS_R;
      p.units[j] = (long long)p.job[j].nb * p.nstrips[j] * p.nchunks[j];
    } else {
      p.units[j] = 0;
      p.nstrips[j] = p.nchunks[j] = 1;
    }
    p.total_units += p.units[j];
  }
  if (p.total_units == 0) return LL_OK;
  static thread_local bool attr_set[64] = {false};
  int dev = 0;
  LL_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    LL_CUDA_OK(cudaFuncSetAttribute(lift_step_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    LL_CUDA_OK(cudaFuncSetAttribute(lift_step_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    LL_CUDA_OK(cudaFuncSetAttribute(lift_step_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    LL_CUDA_OK(cudaFuncSetAttribute(lift_step_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attr_set[dev] = true;
  }
  long long grid = sm_count_cached();
  if (grid > p.total_units) grid = p.total_units;
  const bool f16 = p.f16 && !p.linear;       // without tanh the activations are unbounded: fp16 operands are not safe
  const bool dbg = p.dbg || p.dbg_buf;
  if (f16) {
    if (dbg) lift_step_tc_kernel<true, true><<<(unsigned)grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(p);
    else lift_step_tc_kernel<false, true><<<(unsigned)grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(p);
  } else if (dbg) lift_step_tc_kernel<true, false><<<(unsigned)grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(p);
  else lift_step_tc_kernel<false, false><<<(unsigned)grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(p);
  LL_LAUNCH_OK("lift_step_tc_kernel");
  return LL_OK;
}

int launch_pack_lift_tc(const float* w2, const float* w3, const float* w4, float* blob, cudaStream_t stream) {
  pack_lift_tc16_conv4_kernel<<<8, 256, 0, stream>>>(w4, blob);
  LL_LAUNCH_OK("pack_lift_tc16_conv4_kernel");
  pack_lift_tc_kernel<<<100, 256, 0, stream>>>(w2, w3, blob);
  LL_LAUNCH_OK("pack_lift_tc_kernel");
  pack_lift_tc16_kernel<<<50, 256, 0, stream>>>(w2, w3, blob);
  LL_LAUNCH_OK("pack_lift_tc16_kernel");
  return LL_OK;
}

}  // namespace ll

// rans.cu -- parallel entropy coding of the quantised subbands (SURVEY.md 8f "next #3").
//
// The reference only has a serial, per-coefficient Python coder, and only for its autoregressive model
// (LiftingBasedDWT_net.py:374-556, compress_ar / decompress_ar around compressai.ans).  The entropy layers whose
// contexts depend on *already decoded samples of other levels / phases only* (factorized :182-231, onlyEZWT :759-840,
// ZTBlock :558-757 with its four 2x2 phases) can be decoded a whole subband (phase) at a time, so their symbols are coded
// here with interleaved rANS streams, one warp per stream (the lanes share the CDF evaluations, see the kernels):
//   * state 32 bit, renormalisation by 16-bit words, probabilities quantised to 2^16 (ryg_rans word variant);
//   * image b of a (B, C, hw) tensor owns S streams; stream s codes samples s, s+S, s+2S, ... of the image (neighbouring
//     streams touch neighbouring samples, so their sectors are shared in L2) and every image's bytes are separable;
//   * the distribution of a sample is never tabulated: the coder evaluates the model's own CDF -- the Gaussian of
//     GaussianConditional (sigma clamped at 0.11, mean mu; compressai 1.2.1 entropy_models.py) or the factorized
//     logistic-mixture CDF of EntropyBottleneck -- at the symbol's edges, with the same device code in the encoder and
//     the decoder (bit-identical on the same architecture):
//         C(a) = 2a + floor(F(a - K - 1/2) * (2^16 - 2(2K+2))),   a = k + K in [0, 2K], a = 2K+1 escape
//     (the 2a term guarantees every symbol a frequency >= 1 even if F glitches by an ulp; K = clamp(ceil(6 sigma),
//     15, 2047) for the Gaussian, 255 for the factorized model); |k| > K is an escape followed by 16 raw bits.
// For the ZTBlock layer, which decodes and conditions on plain round(x), the same Gaussian is discretised on the
// integer grid instead (GaussGridDist).
// The symbol is k = round(y - mu) of the *dequantised* value y = round(x - mu) + mu the forward pass returns, and the
// decoder returns k + mu computed the same way, so decode(encode(y)) == y bit for bit.
#include <stdint.h>

#include "ll_common.cuh"

namespace ll {

constexpr int RN_THREADS = 128;
constexpr uint32_t RN_L = 1u << 16;
constexpr int RN_EB_BLOB = 64;   // == EB_BLOB of rate.cu (ll_pack_eb)

// EntropyBottleneck logits (same arithmetic as rate.cu::eb_logits; blob layout of ll_pack_eb)
__device__ __forceinline__ float rn_eb_logits(const float* __restrict__ w, float v) {
  float h[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float t = __fadd_rn(__fmul_rn(w[i], v), w[3 + i]);
    h[i] = __fadd_rn(t, __fmul_rn(w[6 + i], tanhf(t)));
  }
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const float* q = w + 9 + 15 * l;
    float g[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float t = __fmul_rn(q[3 * i], h[0]);
      t = fmaf(q[3 * i + 1], h[1], t);
      t = fmaf(q[3 * i + 2], h[2], t);
      t = __fadd_rn(t, q[9 + i]);
      g[i] = __fadd_rn(t, __fmul_rn(q[12 + i], tanhf(t)));
    }
    h[0] = g[0];
    h[1] = g[1];
    h[2] = g[2];
  }
  float t = __fmul_rn(w[54], h[0]);
  t = fmaf(w[55], h[1], t);
  t = fmaf(w[56], h[2], t);
  return __fadd_rn(t, w[57]);
}

// A sample's distribution: cumulative frequency of alphabet index a in [0, 2K+2] (2K+2 -> 2^16).
struct GaussDist {
  float s, mu;
  int K;
  __device__ __forceinline__ GaussDist(float sigma, float mean) : s(fmaxf(sigma, 0.11f)), mu(mean) {
    const int k = (int)ceilf(6.f * s);
    K = k < 15 ? 15 : (k > 2047 ? 2047 : k);
  }
  __device__ __forceinline__ float centre() const { return mu; }
  __device__ __forceinline__ uint32_t C(int a) const {
    if (a <= 0) return 0u;
    if (a >= 2 * K + 2) return 65536u;
    const uint32_t M = 65536u - 2u * (2 * K + 2);
    const float t = (float)(a - K) - 0.5f;
    const float F = 0.5f * erfcf(-0.70710678118654752440f * __fdiv_rn(t, s));
    const uint32_t q = (uint32_t)fminf(floorf(F * (float)M), (float)M);
    return 2u * a + q;
  }
};

// Gaussian N(mu, sigma) discretised on the INTEGER grid: the symbol is the plain round(x) that
// DWTConditioned2EntropyLayerZTBlock decodes and conditions on (:719-724,754), coded relative to c = round(mu);
// P(k) = Phi((c + k + 1/2 - mu) / sigma) - Phi((c + k - 1/2 - mu) / sigma).
struct GaussGridDist {
  float s, c, delta;
  int K;
  __device__ __forceinline__ GaussGridDist(float sigma, float mean) : s(fmaxf(sigma, 0.11f)), c(rintf(mean)) {
    delta = __fsub_rn(c, mean);
    const int k = (int)ceilf(6.f * s) + 1;
    K = k < 15 ? 15 : (k > 2047 ? 2047 : k);
  }
  __device__ __forceinline__ float centre() const { return c; }
  __device__ __forceinline__ uint32_t C(int a) const {
    if (a <= 0) return 0u;
    if (a >= 2 * K + 2) return 65536u;
    const uint32_t M = 65536u - 2u * (2 * K + 2);
    const float t = __fadd_rn((float)(a - K) - 0.5f, delta);
    const float F = 0.5f * erfcf(-0.70710678118654752440f * __fdiv_rn(t, s));
    const uint32_t q = (uint32_t)fminf(floorf(F * (float)M), (float)M);
    return 2u * a + q;
  }
};

struct EbDist {
  const float* w;
  float med;
  int K;
  __device__ __forceinline__ EbDist(const float* blob_c) : w(blob_c), med(blob_c[58]), K(255) {}
  __device__ __forceinline__ float centre() const { return med; }
  __device__ __forceinline__ uint32_t C(int a) const {
    if (a <= 0) return 0u;
    if (a >= 2 * K + 2) return 65536u;
    const uint32_t M = 65536u - 2u * (2 * K + 2);
    const float t = __fadd_rn(med, (float)(a - K) - 0.5f);
    const float F = 1.f / (1.f + expf(-rn_eb_logits(w, t)));
    const uint32_t q = (uint32_t)fminf(floorf(F * (float)M), (float)M);
    return 2u * a + q;
  }
};

struct RansEnc {   // writes 16-bit words backwards from the end of the stream's scratch region
  uint32_t x;
  uint16_t* buf;
  int pos;
  bool writer;   // every lane of the stream's warp runs the state machine, one of them stores the words
  __device__ __forceinline__ RansEnc(uint16_t* b, int cap, bool w) : x(RN_L), buf(b), pos(cap), writer(w) {}
  __device__ __forceinline__ void put(uint32_t start, uint32_t freq) {
    if (x >= (freq << 16)) {
      --pos;
      if (writer) buf[pos] = (uint16_t)(x & 0xffffu);
      x >>= 16;
    }
    x = ((x / freq) << 16) + (x % freq) + start;
  }
  __device__ __forceinline__ int finish(int cap) {
    pos -= 2;
    if (writer) {
      buf[pos + 1] = (uint16_t)(x & 0xffffu);
      buf[pos] = (uint16_t)(x >> 16);
    }
    return cap - pos;
  }
};

struct RansDec {
  uint32_t x;
  const uint16_t* buf;
  __device__ __forceinline__ RansDec(const uint16_t* b) : x(((uint32_t)b[0] << 16) | b[1]), buf(b + 2) {}
  __device__ __forceinline__ uint32_t slot() const { return x & 0xffffu; }
  __device__ __forceinline__ void advance(uint32_t start, uint32_t freq) {
    x = freq * (x >> 16) + (x & 0xffffu) - start;
    if (x < RN_L) x = (x << 16) | *buf++;
  }
};

// (start, freq) of a sample under its distribution -- independent of the coder state, so the warp evaluates 32 samples
// at once.  Escape: ``raw`` is pushed (frequency 1) before the escape symbol, because the decoder pops the escape first.
struct EncSym {
  uint32_t start, freq, raw;
  bool esc;
};
template <class D>
__device__ __forceinline__ EncSym enc_symbol(const D& d, float y) {
  EncSym r;
  const float kf = rintf(__fsub_rn(y, d.centre()));
  const int K = d.K;
  r.esc = !(fabsf(kf) <= (float)K);
  r.raw = 0u;
  int a = (int)kf + K;
  if (r.esc) {
    const float kc = fminf(fmaxf(kf, -32768.f), 32767.f);
    r.raw = (uint32_t)((int)kc + 32768);
    a = 2 * K + 1;
  }
  r.start = d.C(a);
  r.freq = d.C(a + 1) - r.start;      // C(2K + 2) = 2^16
  return r;
}

template <class D>
__device__ __forceinline__ float dec_sample(RansDec& r, const D& d, int lane) {
  // Wanted: the largest a in [0, 2K+1] with C(a) <= slot.  Every C() is an erfc (or logistic-mixture) evaluation on
  // the stream's serial dependency chain, so the WARP that owns the stream searches cooperatively: each round the 32
  // lanes probe 32 evenly spaced candidates at once and a ballot keeps the sub-interval that holds the answer --
  // 1 round for the 32-symbol alphabets of small sigma, 3 for the widest (4096 symbols) instead of 7 .. 14 dependent
  // evaluations of a bisection.  C(lo) and C(hi + 1) ride along, so no evaluation is repeated.  All state is
  // warp-uniform (every lane runs the same rANS state machine).
  const uint32_t sl = r.slot();
  const int K = d.K, top = 2 * K + 1;
  int lo = 0, hi = top;
  uint32_t c_lo = 0u, c_up = 65536u;     // C(lo), C(hi + 1)
  while (lo < hi) {
    const int stepw = (hi - lo + 31) >> 5;
    const int b = lo + (lane + 1) * stepw;
    const uint32_t cb = b <= hi ? d.C(b) : 65536u;
    const unsigned le = __ballot_sync(0xffffffffu, b <= hi && cb <= sl);   // C is increasing: a prefix of the lanes
    const int cnt = __popc(le);
    const uint32_t c_in = __shfl_sync(0xffffffffu, cb, cnt > 0 ? cnt - 1 : 0);
    const uint32_t c_out = __shfl_sync(0xffffffffu, cb, cnt < 32 ? cnt : 31);
    if (cnt < 32 && lo + (cnt + 1) * stepw <= hi) {
      hi = lo + (cnt + 1) * stepw - 1;
      c_up = c_out;
    }
    if (cnt > 0) {
      lo += cnt * stepw;
      c_lo = c_in;
    }
  }
  r.advance(c_lo, c_up - c_lo);
  float kf;
  if (lo <= 2 * K) kf = (float)(lo - K);
  else {
    const uint32_t v = r.slot();
    r.advance(v, 1u);
    kf = (float)((int)v - 32768);
  }
  return __fadd_rn(kf, d.centre());
}

// MODE 0: Gaussian centred on mu, ms (B, 2C, hw) with channel 2c = sigma, 2c+1 = mu.  MODE 1: factorized, blob (C, 64).
// MODE 2: Gaussian on the integer grid (same ms layout).
template <int MODE>
__global__ void __launch_bounds__(RN_THREADS) rans_encode_kernel(const float* __restrict__ y, const float* __restrict__ par, int B,
                                                                int C, long long hw, int S, uint16_t* __restrict__ scratch,
                                                                int cap, int* __restrict__ counts) {
  // One WARP per stream.  The stream's samples s + j S are coded backwards (j = ns-1 .. 0), 32 at a time: lane l
  // evaluates the CDF edges of sample j0 - l (the erfc work, which does not depend on the coder state, runs 32-wide),
  // then every lane replays the 32 (start, freq) pairs through the serial rANS state machine (a compare, a 32-bit
  // divide and an occasional 16-bit store per symbol).  The next chunk's loads are issued before the serial part.
  const long long st = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (st >= (long long)B * S) return;
  const long long b = st / S;
  const int s = (int)(st % S);
  const long long N = (long long)C * hw;
  const float* yb = y + b * N;
  const float* pb = par + b * 2 * C * hw;
  RansEnc enc(scratch + st * cap, cap, lane == 0);
  const long long ns = s < N ? (N - s + S - 1) / S : 0;
  float v = 0.f, sg = 1.f, mu = 0.f;
  int ch = 0;
  auto fetch = [&](long long j) {
    if (j < 0) return;
    const long long e = s + j * S;
    ch = (int)(e / hw);
    const long long pix = e - ch * hw;
    v = yb[e];
    if (MODE != 1) {
      sg = pb[2 * ch * hw + pix];
      mu = pb[(2 * ch + 1) * hw + pix];
    }
  };
  fetch(ns - 1 - lane);
  for (long long j0 = ns - 1; j0 >= 0; j0 -= 32) {
    EncSym sym = {0u, 1u, 0u, false};
    if (j0 - lane >= 0) {
      if (MODE == 0) sym = enc_symbol(GaussDist(sg, mu), v);
      else if (MODE == 2) sym = enc_symbol(GaussGridDist(sg, mu), v);
      else sym = enc_symbol(EbDist(par + (size_t)ch * RN_EB_BLOB), v);
    }
    fetch(j0 - 32 - lane);
    const int n = j0 + 1 < 32 ? (int)(j0 + 1) : 32;
    for (int i = 0; i < n; ++i) {
      const uint32_t st_i = __shfl_sync(0xffffffffu, sym.start, i);
      const uint32_t fr_i = __shfl_sync(0xffffffffu, sym.freq, i);
      const uint32_t raw_i = __shfl_sync(0xffffffffu, sym.raw, i);
      const bool esc_i = __shfl_sync(0xffffffffu, (int)sym.esc, i) != 0;
      if (esc_i) enc.put(raw_i, 1u);
      enc.put(st_i, fr_i);
    }
  }
  const int words = enc.finish(cap);
  if (lane == 0) counts[st] = words;
}

__global__ void rans_pack_kernel(const uint16_t* __restrict__ scratch, const int* __restrict__ counts, const long long* __restrict__ offsets,
                                 long long nstreams, int cap, uint16_t* __restrict__ packed) {
  // one warp per stream
  const long long st = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (st >= nstreams) return;
  const int n = counts[st];
  const uint16_t* src = scratch + st * cap + (cap - n);
  uint16_t* dst = packed + offsets[st];
  for (int i = lane; i < n; i += 32) dst[i] = src[i];
}

template <int MODE>
__global__ void __launch_bounds__(RN_THREADS) rans_decode_kernel(const uint16_t* __restrict__ packed, const long long* __restrict__ offsets,
                                                                const float* __restrict__ par, int B, int C, long long hw, int S,
                                                                float* __restrict__ y) {
  // one WARP per stream (see dec_sample); lane 0 writes the decoded samples
  const long long st = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (st >= (long long)B * S) return;
  const long long b = st / S;
  const int s = (int)(st % S);
  const long long N = (long long)C * hw;
  float* yb = y + b * N;
  RansDec dec(packed + offsets[st]);
  // same software pipeline as the encoder: parameters of the next sample are in flight while this one is decoded
  int c = s < N ? (int)(s / hw) : 0;
  long long pix = s - c * hw;
  const float* pb = par + b * 2 * C * hw;
  float sg = 1.f, mu = 0.f;
  if (MODE != 1 && s < N) {
    sg = pb[2 * c * hw + pix];
    mu = pb[(2 * c + 1) * hw + pix];
  }
  for (long long e = s; e < N; e += S) {
    const int cc = c;
    const float sgc = sg, muc = mu;
    if (e + S < N) {
      pix += S;
      while (pix >= hw) {
        pix -= hw;
        ++c;
      }
      if (MODE != 1) {
        sg = pb[2 * c * hw + pix];
        mu = pb[(2 * c + 1) * hw + pix];
      }
    }
    float out;
    if (MODE == 0) out = dec_sample(dec, GaussDist(sgc, muc), lane);
    else if (MODE == 2) out = dec_sample(dec, GaussGridDist(sgc, muc), lane);
    else out = dec_sample(dec, EbDist(par + (size_t)cc * RN_EB_BLOB), lane);
    if (lane == 0) yb[e] = out;
  }
}

}  // namespace ll

using namespace ll;

extern "C" {

int64_t ll_rans_stream_cap(int64_t n_per_image, int S) {
  if (n_per_image < 0 || S <= 0) return -1;
  const int64_t per = (n_per_image + S - 1) / S;
  return 2 * per + 4;   // <= 2 words per sample (escape + raw), 2 words of final state, slack
}

static int rans_args(const char* who, const void* a, const void* b, const void* c, const void* d, int B, int C, int64_t hw, int S) {
  if (B < 0 || C <= 0 || hw < 0 || S <= 0) return fail(LL_EINVAL, "%s: bad extents", who);
  if ((long long)B * hw > 0 && (!a || !b || !c || !d)) return fail(LL_EINVAL, "%s: null pointer", who);
  if ((long long)B * S > 0x7fffffffLL) return fail(LL_EINVAL, "%s: too many streams", who);
  return LL_OK;
}

int ll_rans_encode(int mode, const float* y, const float* par, int B, int C, int64_t hw, int S, uint16_t* scratch,
                   int32_t* counts, ll_stream_t stream) {
  int rc = rans_args("ll_rans_encode", y, par, scratch, counts, B, C, hw, S);
  if (rc) return rc;
  if (mode < 0 || mode > 2) return fail(LL_EINVAL, "ll_rans_encode: mode must be 0 (gaussian), 1 (factorized) or 2 (gaussian, integer grid)");
  const long long nst = (long long)B * S;
  if (nst == 0 || hw == 0) return LL_OK;
  const int cap = (int)ll_rans_stream_cap((int64_t)C * hw, S);
  const unsigned blocks = (unsigned)((nst * 32 + RN_THREADS - 1) / RN_THREADS);   // one warp per stream
  if (mode == 0) rans_encode_kernel<0><<<blocks, RN_THREADS, 0, as_stream(stream)>>>(y, par, B, C, hw, S, scratch, cap, counts);
  else if (mode == 2) rans_encode_kernel<2><<<blocks, RN_THREADS, 0, as_stream(stream)>>>(y, par, B, C, hw, S, scratch, cap, counts);
  else rans_encode_kernel<1><<<blocks, RN_THREADS, 0, as_stream(stream)>>>(y, par, B, C, hw, S, scratch, cap, counts);
  LL_LAUNCH_OK("rans_encode_kernel");
  return LL_OK;
}

int ll_rans_pack(const uint16_t* scratch, const int32_t* counts, const int64_t* offsets, int64_t nstreams, int cap,
                 uint16_t* packed, ll_stream_t stream) {
  if (nstreams < 0 || cap <= 0) return fail(LL_EINVAL, "ll_rans_pack: bad extents");
  if (nstreams == 0) return LL_OK;
  if (!scratch || !counts || !offsets || !packed) return fail(LL_EINVAL, "ll_rans_pack: null pointer");
  const unsigned blocks = (unsigned)((nstreams * 32 + 255) / 256);
  rans_pack_kernel<<<blocks, 256, 0, as_stream(stream)>>>(scratch, counts, reinterpret_cast<const long long*>(offsets), nstreams, cap, packed);
  LL_LAUNCH_OK("rans_pack_kernel");
  return LL_OK;
}

int ll_rans_decode(int mode, const uint16_t* packed, const int64_t* offsets, const float* par, int B, int C, int64_t hw, int S,
                   float* y, ll_stream_t stream) {
  int rc = rans_args("ll_rans_decode", packed, offsets, par, y, B, C, hw, S);
  if (rc) return rc;
  if (mode < 0 || mode > 2) return fail(LL_EINVAL, "ll_rans_decode: mode must be 0 (gaussian), 1 (factorized) or 2 (gaussian, integer grid)");
  const long long nst = (long long)B * S;
  if (nst == 0 || hw == 0) return LL_OK;
  const unsigned blocks = (unsigned)((nst * 32 + RN_THREADS - 1) / RN_THREADS);   // one warp per stream
  if (mode == 0) rans_decode_kernel<0><<<blocks, RN_THREADS, 0, as_stream(stream)>>>(packed, reinterpret_cast<const long long*>(offsets), par, B, C, hw, S, y);
  else if (mode == 2) rans_decode_kernel<2><<<blocks, RN_THREADS, 0, as_stream(stream)>>>(packed, reinterpret_cast<const long long*>(offsets), par, B, C, hw, S, y);
  else rans_decode_kernel<1><<<blocks, RN_THREADS, 0, as_stream(stream)>>>(packed, reinterpret_cast<const long long*>(offsets), par, B, C, hw, S, y);
  LL_LAUNCH_OK("rans_decode_kernel");
  return LL_OK;
}

}  // extern "C"

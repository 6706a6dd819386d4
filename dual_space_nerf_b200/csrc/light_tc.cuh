// Shading on the tensor cores: normal_local2world + LightingMLP (model/spacenet.py:165-188, 278-298).
//
// Tile = 128 active samples per CTA iteration, one thread per sample (4 warps):
//   A. world normal (through the nearest canonical triangle found by canon_nearest_kernel beforehand), world position,
//      view direction (shade_inputs); the 9 -> 128 first layer runs on the tensor core as well: the nine inputs are split
//      into fp16 hi + lo and laid out as one K = 32 operand row [in_hi | in_lo | in_hi | 1 | 1 | 0 0 0] against
//      [W1_hi ; W1_hi ; W1_lo ; b_hi ; b_lo] (= the 3-pass split of mlp_tc.cuh folded into K, bias included; two MMAs).  In
//      fp32 registers this layer was 1152 FMAs + 320 shared-memory loads per sample, two thirds of the kernel's instructions
//      (ncu: 154 M warp instructions, issue-bound at 16 warps per SM).  Its accumulator is read back, ReLU'd and written as
//      the fp16 A operand of the second layer (K-major core matrices) in place of the first operand;
//   B. one elected lane issues 8 tcgen05.mma (M = 128, N = 128, K = 16; A and B from shared memory): the 128 x 128
//      second layer.  Its weights are packed on the host into the B-operand image and fetched once per CTA with
//      cp.async.bulk, so they stay resident in shared memory for every tile of the CTA;
//   C. tcgen05.ld of the thread's accumulator row, bias + ReLU + 128 -> 1 dot + ELU in fp32, colour = (out+1)*essence.
// A CTA runs TWO such tiles side by side (256 threads: two independent 4-warp halves with their own A operand,
// accumulator, mbarrier and named barrier) on one copy of the second-layer weights: ~103 KB of shared memory and 256 TMEM
// columns per CTA, two CTAs = 16 warps per SM (registers and TMEM allow no more).  With so few warps the kernel cannot hide
// chains of dependent gathers: the nearest-centroid search used to run inside shade_inputs here (1.42 ms); as a separate
// full-occupancy kernel it takes 0.31 ms and this kernel 0.36 ms.  Only the middle layer is rounded to fp16 (single pass): the lighting term is smooth and enters the
// colour linearly (measured effect on |d rgb| is ~1e-5); the fp32 SIMT kernel in shade.cuh remains as
// the verification path.
#pragma once
#include "mlp_tc.cuh"
#include "shade.cuh"

namespace dsn {

constexpr int LT_THREADS = 256;
constexpr int LT_ROWS = 128;                        // rows (samples) per half
// shared memory map; PARTS = 1 (hi only) or 2 (hi + lo) copies of every fp16 operand
template <int PARTS> struct LtMap {
  static constexpr uint32_t W2 = 0;                         // B operand: PARTS x [16 k-chunks][128 rows][8] fp16 = PARTS x 32 KB
  static constexpr uint32_t A = PARTS * 32768;              // A operands: 2 halves x PARTS x 32 KB
  static constexpr uint32_t A_HALF = PARTS * 32768;         // bytes per half
  static constexpr uint32_t W1 = A + 2 * A_HALF;            // first layer as a B operand: [4 k-chunks][128 rows][8] fp16 = 8 KB
  static constexpr uint32_t B2 = W1 + 8192;
  static constexpr uint32_t W3 = B2 + 512;
  static constexpr uint32_t BAR = W3 + 512;                 // 3 mbarriers + tmem slot
  static constexpr uint32_t SMEM = BAR + 64;
};
constexpr uint32_t LT_SMEM = LtMap<1>::SMEM;
constexpr uint32_t LT_SMEM3 = LtMap<2>::SMEM;
constexpr uint32_t LT_TMEM_COLS = 256;

template <bool P3>
__global__ void __launch_bounds__(LT_THREADS, P3 ? 1 : 2) light_tc_kernel(ShadeArgs a, LightWeights L, const uint8_t* __restrict__ w2_packed, Grid gc) {
  using M = LtMap<P3 ? 2 : 1>;
  constexpr uint32_t LT_SM_W2 = M::W2, LT_SM_A = M::A, LT_SM_W1 = M::W1, LT_SM_B2 = M::B2, LT_SM_W3 = M::W3, LT_SM_BAR = M::BAR;
  constexpr uint32_t W2_BYTES = (P3 ? 2u : 1u) * 32768u;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int half = threadIdx.x >> 7;                 // independent 4-warp half of the CTA
  const int row = threadIdx.x & (LT_ROWS - 1);
  const int hwarp = (threadIdx.x >> 5) & 3;          // warp within the half = TMEM lane quarter
  const uint32_t bar_w = sbase + LT_SM_BAR, bar_mma = bar_w + 8 + 8 * half;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + LT_SM_BAR + 24);
  float* b2 = reinterpret_cast<float*>(smem + LT_SM_B2);
  float* w3 = reinterpret_cast<float*>(smem + LT_SM_W3);
  uint8_t* a_op = smem + LT_SM_A + half * M::A_HALF;  // hi part; the lo part (P3) follows 32 KB later
  auto half_bar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory"); };

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_w + 8, 1);
    mbar_init(bar_w + 16, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(LT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 128; i += LT_THREADS) { b2[i] = L.b2[i]; w3[i] = L.w3[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot + (uint32_t)half * 128u;
  if (threadIdx.x == 0) {  // second-layer weights: one bulk copy, resident for the whole kernel
    mbar_expect_tx(bar_w, W2_BYTES + 8192u);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sbase + LT_SM_W2), "l"(w2_packed), "r"(W2_BYTES), "r"(bar_w) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sbase + LT_SM_W1), "l"(w2_packed + 65536), "r"(8192u), "r"(bar_w) : "memory");  // first layer follows W2 hi + lo
  }
  const int64_t n_active = a.n_active ? (int64_t)*a.n_active : a.n_active_host;
  const int64_t n_tiles = (n_active + LT_ROWS - 1) / LT_ROWS;
  const uint32_t t_lane = tmem + ((uint32_t)(hwarp * 32) << 16);
  uint32_t mma_phase = 0;
  bool w_ready = false;
  // the records of the next tile are fetched while the current one is processed (they head a chain of ~6 dependent gathers)
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ac_n = zero4, ma_n = zero4, mg_n = zero4;
  int ci_n = -1;
  {
    const int64_t t0 = ((int64_t)blockIdx.x * 2 + half) * LT_ROWS + row;
    if (t0 < n_active) { ac_n = a.active[t0]; ma_n = a.mlp_a[t0]; mg_n = a.mlp_g[t0]; if (a.active_cidx) ci_n = a.active_cidx[t0]; }
  }
  for (int64_t tile = (int64_t)blockIdx.x * 2 + half; tile < n_tiles; tile += (int64_t)gridDim.x * 2) {
    const int64_t t = tile * LT_ROWS + row;
    const bool live = t < n_active;
    float in[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float4 ac = ac_n, ma = ma_n, mg = mg_n;
    const int ci = ci_n;
    {
      const int64_t tn = t + (int64_t)gridDim.x * 2 * LT_ROWS;
      if (tn < n_active) { ac_n = a.active[tn]; ma_n = a.mlp_a[tn]; mg_n = a.mlp_g[tn]; if (a.active_cidx) ci_n = a.active_cidx[tn]; }
    }
    int sample = 0;
    if (live) shade_inputs(a, gc, ac, mg, in, sample, ci);
    // ---- first layer on the tensor core: operand row [in_hi (9) | in_lo (9) | in_hi (9) | 1 | 1 | 0 0 0] -> chunks 0..3 of a_op
    {
      __half hh[9], hl[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        hh[k] = __float2half_rn(in[k]);
        hl[k] = __float2half_rn(in[k] - __half2float(hh[k]));
      }
      const __half one = __float2half_rn(1.0f), zero = __float2half_rn(0.0f);
      auto pk2 = [](__half x, __half y) -> uint32_t { __half2 h = __halves2half2(x, y); return *reinterpret_cast<uint32_t*>(&h); };
      const uint4 c0 = make_uint4(pk2(hh[0], hh[1]), pk2(hh[2], hh[3]), pk2(hh[4], hh[5]), pk2(hh[6], hh[7]));
      const uint4 c1 = make_uint4(pk2(hh[8], hl[0]), pk2(hl[1], hl[2]), pk2(hl[3], hl[4]), pk2(hl[5], hl[6]));
      const uint4 c2 = make_uint4(pk2(hl[7], hl[8]), pk2(hh[0], hh[1]), pk2(hh[2], hh[3]), pk2(hh[4], hh[5]));
      const uint4 c3 = make_uint4(pk2(hh[6], hh[7]), pk2(hh[8], one), pk2(one, zero), pk2(zero, zero));
      *reinterpret_cast<uint4*>(a_op + 0u * (LT_ROWS * 16) + row * 16) = c0;
      *reinterpret_cast<uint4*>(a_op + 1u * (LT_ROWS * 16) + row * 16) = c1;
      *reinterpret_cast<uint4*>(a_op + 2u * (LT_ROWS * 16) + row * 16) = c2;
      *reinterpret_cast<uint4*>(a_op + 3u * (LT_ROWS * 16) + row * 16) = c3;
    }
    fence_proxy_async();
    half_bar();
    constexpr uint32_t DHI = (128u >> 4) | (1u << 14);
    constexpr uint32_t LBO = ((LT_ROWS * 16) >> 4) << 16;
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((128u >> 4) << 24);
    if (hwarp == 0) {
      if (!w_ready) { mbar_wait(bar_w, 0); w_ready = true; }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a0 = LBO | ((sbase + LT_SM_A + half * M::A_HALF) >> 4), b1 = LBO | ((sbase + LT_SM_W1) >> 4);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint32_t step = (uint32_t)k * ((2 * LT_ROWS * 16) >> 4);
          tc_mma_ss(tmem, ((uint64_t)DHI << 32) | (a0 + step), ((uint64_t)DHI << 32) | (b1 + step), IDESC, k > 0 ? 1u : 0u);
        }
        tc_commit(bar_mma);
      }
      __syncwarp();
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    // ---- ReLU of the first layer -> fp16 A operand of the second (its hi part overwrites the first operand: the MMAs that read it are done)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(t_lane + c * 32, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t pk[4], pl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float h0 = fmaxf(__uint_as_float(v[q * 8 + 2 * e]), 0.f), h1 = fmaxf(__uint_as_float(v[q * 8 + 2 * e + 1]), 0.f);
          pk[e] = pack_h2(h0, h1);
          if (P3) {
            const float2 hf = __half22float2(as_h2(pk[e]));
            pl[e] = pack_h2(h0 - hf.x, h1 - hf.y);
          }
        }
        const uint32_t kc = (uint32_t)(c * 4 + q);
        *reinterpret_cast<uint4*>(a_op + kc * (LT_ROWS * 16) + row * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        if (P3) *reinterpret_cast<uint4*>(a_op + 32768 + kc * (LT_ROWS * 16) + row * 16) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
      }
    }
    tc_fence_before();
    fence_proxy_async();
    half_bar();  // every thread has read its accumulator row and written its operand row
    // ---- second layer on the tensor core
    if (hwarp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a0 = LBO | ((sbase + LT_SM_A + half * M::A_HALF) >> 4), b0 = LBO | ((sbase + LT_SM_W2) >> 4);
        constexpr uint32_t LO = 32768u >> 4;  // hi -> lo part of either operand
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t step = (uint32_t)k * ((2 * LT_ROWS * 16) >> 4);
          tc_mma_ss(tmem, ((uint64_t)DHI << 32) | (a0 + step), ((uint64_t)DHI << 32) | (b0 + step), IDESC, k > 0 ? 1u : 0u);
          if (P3) {  // x_hi * w_lo + x_lo * w_hi
            tc_mma_ss(tmem, ((uint64_t)DHI << 32) | (a0 + step), ((uint64_t)DHI << 32) | (b0 + LO + step), IDESC, 1u);
            tc_mma_ss(tmem, ((uint64_t)DHI << 32) | (a0 + LO + step), ((uint64_t)DHI << 32) | (b0 + step), IDESC, 1u);
          }
        }
        tc_commit(bar_mma);
      }
      __syncwarp();
    }
    mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1;
    tc_fence_after();
    // ---- third layer + ELU (fp32)
    float out = L.b3;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(t_lane + c * 32, v);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 bb = *reinterpret_cast<const float4*>(b2 + c * 32 + i);
        const float4 ww = *reinterpret_cast<const float4*>(w3 + c * 32 + i);
        out = fmaf(fmaxf(__uint_as_float(v[i]) + bb.x, 0.f), ww.x, out);
        out = fmaf(fmaxf(__uint_as_float(v[i + 1]) + bb.y, 0.f), ww.y, out);
        out = fmaf(fmaxf(__uint_as_float(v[i + 2]) + bb.z, 0.f), ww.z, out);
        out = fmaf(fmaxf(__uint_as_float(v[i + 3]) + bb.w, 0.f), ww.w, out);
      }
    }
    const float light = (out > 0.f ? out : expm1f(out)) + 1.0f;
    if (live) a.raw[sample] = make_float4(light * ma.y, light * ma.z, light * ma.w, ma.x);
    tc_fence_before();
    half_bar();  // accumulator and A operand are reused by the half's next tile
    tc_fence_after();
  }
  if (hwarp == 0 && !w_ready) mbar_wait(bar_w, 0);  // never leave with the bulk copy in flight
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(LT_TMEM_COLS) : "memory");
  }
}

// host: pack lights_encoding.2.weight [out=128][in=128] as the B operand image (B[n][k] = W[n][k]): hi part, then lo part;
// then lights_encoding.0 (w1 [128][9], b1 [128]) as the K = 32 B operand [w_hi (9) | w_hi (9) | w_lo (9) | b_hi | b_lo | 0 0 0]
inline void light_pack_w2(const std::vector<float>& w2, const std::vector<float>& w1, const std::vector<float>& b1, std::vector<__half>& out) {
  out.assign((size_t)2 * 128 * 128 + 128 * 32, __float2half_rn(0.f));
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 32; ++k) {
      float v = 0.f;
      auto hi = [](float x) { return __half2float(__float2half_rn(x)); };
      if (k < 9) v = hi(w1[(size_t)n * 9 + k]);
      else if (k < 18) v = hi(w1[(size_t)n * 9 + k - 9]);
      else if (k < 27) v = w1[(size_t)n * 9 + k - 18] - hi(w1[(size_t)n * 9 + k - 18]);
      else if (k == 27) v = hi(b1[n]);
      else if (k == 28) v = b1[n] - hi(b1[n]);
      out[(size_t)2 * 128 * 128 + ((size_t)(k / 8) * 128 + n) * 8 + (k % 8)] = __float2half_rn(v);
    }
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 128; ++k) {
      const float w = w2[(size_t)n * 128 + k];
      const __half hh = __float2half_rn(w);
      const size_t o = ((size_t)(k / 8) * 128 + n) * 8 + (k % 8);
      out[o] = hh;
      out[(size_t)128 * 128 + o] = __float2half_rn(w - __half2float(hh));
    }
}

}  // namespace dsn

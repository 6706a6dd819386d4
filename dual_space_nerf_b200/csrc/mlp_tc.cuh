// SpaceNet forward + analytic density gradient on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// Restates model/spacenet.py:93-148 (SpaceNet.forward: PE -> 4x256 -> [h|PE] -> 3x256 -> density,
// rgb head) and :301-311 (gradient = d density / d xyz_cano, which the reference gets from
// autograd) for a tile of 128 canonical points per CTA (UMMA M = 128, cta_group::1).
//
// Numerics
//   * forward GEMMs run 3 MMAs per k-step (x_hi*w_hi + x_hi*w_lo + x_lo*w_hi, fp16 operands, fp32
//     accumulate in TMEM): single-pass fp16/bf16/tf32 misses the 1e-4 parity bound by >10x because a
//     rounding-sized change of a pre-activation flips ReLUs of the gradient path;
//   * the backward-data chain for the normal runs single-pass fp16 through W^T using the ReLU bits
//     recorded in the forward pass; the seed is w_dens / max|w_dens| so every gradient stays inside
//     fp16 range (the normal is scale invariant);
//   * the rgb head's 256 -> 128 layer runs single-pass fp16 as well: the essence enters the colour linearly and is
//     not amplified by the density head (oracle emulation: max |d rgb| 2.9e-5 vs 2.7e-5 with 3 passes).
//
// Pipeline (CTA pair = two SMs of one TPC, cta_group::2; 18 warps per CTA)
//   * measured on B200 (profiles/): an MMA whose A operand comes from shared memory occupies the tensor pipe for
//     >= 128 cycles whatever N is (the 128 x 16 A tile is read at 32 B/cycle), so every MMA here is N = 256
//     (128 cycles = the ideal rate);
//   * with cta_group::1 the kernel was SHARED-MEMORY-BANDWIDTH bound: per MMA the tensor core reads A (4 KB) and B
//     (8 KB) = 96 B/clk of the SM's 128 B/clk, and the weight stream (TMA writes) plus the epilogue's operand stores
//     add ~60 B/clk; MMAs issued every ~154 cycles.  cta_group::2 pairs two SMs on one M = 256 tile (each CTA owns
//     128 points and its own accumulators): every CTA holds and streams only HALF of each weight slab (its N/2 rows
//     of B), the tensor cores fetch the other half from the peer's shared memory, so operand reads drop to 64 B/clk,
//     the L2 -> SM weight traffic halves and the ring doubles its depth (8 x 8 KB) in the same footprint;
//   * the leader CTA's MMA warp issues for both; the peer's warp relays "my half landed" to the leader; epilogue warps
//     of both CTAs arrive (remotely for the peer) on the leader's hand-off barriers; tcgen05.commit multicasts;
//   * the A operand (fp16 hi + lo, K-major no-swizzle core matrices, 128 KB) is rewritten in place by the
//     epilogue; two 256-column TMEM accumulators alternate between layers, so the epilogue of layer l overlaps
//     the MMAs of layer l+1: the epilogue publishes its output in four 64-column quarters and the next layer's
//     k-steps start on each quarter as it lands;
//   * 16 epilogue warps (4 per TMEM lane quarter, 16 columns of every 64-column quarter each): the epilogue is
//     latency bound (tcgen05.ld, fences), so it is spread over 4 warps per scheduler, the next quarter's
//     tcgen05.ld is in flight while the current one is processed, the arithmetic is packed (fp32x2 add/fma,
//     half2 compare for the ReLU bits) and one lane per warp arrives on the hand-off barriers;
//   * layer 4's positional-encoding k-steps do not depend on layer 3's epilogue and are issued first;
//   * weights are pre-packed on the host into the exact shared-memory image, in consumption order, and streamed
//     by one elected lane with cp.async.bulk (TMA engine) into a 4x16 KB mbarrier ring;
//   * one elected lane of a converged warp issues all tcgen05.mma (slab-templated: descriptors are
//     "base + immediate"); tcgen05.commit releases ring slots and publishes finished accumulators;
//   * ReLU bits live in registers (14 words per thread), d sigma / d PE of layer 4 is parked in the idle A-lo
//     region during the backward chain, cross-thread partial sums go through shared memory once per tile.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace dsn {

constexpr int TC_TILE = 128;
constexpr int TC_EPI_WARPS = 16;
constexpr int TC_SUBS = TC_EPI_WARPS / 4;          // column sub-blocks per 64-column quarter
constexpr int TC_CPT = 64 / TC_SUBS;               // columns per thread per quarter (16)
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32;
constexpr int TC_STAGES = 4;
constexpr uint32_t TC_STAGE_BYTES = 16384;  // per CTA: half of a weight slab
// shared memory map (bytes)
constexpr uint32_t SM_A_HI = 0;          // A operand, fp16 hi: [K/8 = 32 chunks][128 rows][8]
constexpr uint32_t SM_A_LO = 65536;      // A operand, fp16 lo (forward only); backward: fp32 stash of layer 4's d sigma / d PE + exchange
constexpr uint32_t SM_PE_HI = 131072;    // positional encoding hi, 8 chunks
constexpr uint32_t SM_PE_LO = 147456;    // positional encoding lo
constexpr uint32_t SM_RING = 163840;
constexpr uint32_t SM_BAR = SM_RING + TC_STAGES * TC_STAGE_BYTES;  // 229376
constexpr uint32_t SM_RGBW = SM_BAR + 256;             // fp32: b_rgb1 [128], w_rgb2 [3][128] (rgb tail of every tile)
constexpr uint32_t TC_SMEM = SM_RGBW + 2048;
constexpr uint32_t A_CHUNK = TC_TILE * 16;  // bytes between consecutive 8-wide K chunks of an A operand
constexpr uint32_t SM_STASH = SM_A_LO;           // [64 PE columns][128 rows] fp32 (32 KB), backward chain only
constexpr uint32_t SM_XCH = SM_A_LO + 32768;     // [3 subs][8][128 rows] fp32 partial sums at the end of a tile
// tensor memory map (32-bit columns)
constexpr uint32_t TM_ACC = 256;    // accumulator b (= op & 1) starts at column b * TM_ACC
constexpr uint32_t TM_COLS = 512;

constexpr int TC_NUM_OPS = 15;
enum { A_ACT = 0, A_PE = 1, A_PE_ACT = 2, A_ACT_LO = 3 };  // A_PE_ACT: k-steps 0..3 come from the PE region, the rest from the activations;
                                                            // A_ACT_LO: operand parked in the A-lo region (seed of the backward chain)

enum { K_FWD = 0, K_RGB = 1, K_BWD = 2, K_BW4 = 3, K_BW0 = 4, K_RGB3 = 5 };

struct TcOp {
  uint32_t src_off;     // byte offset of the op's first slab in the packed weight blob
  uint32_t slab_bytes;  // bytes per slab HALF (one CTA's share, <= TC_STAGE_BYTES, multiple of 32); a slab = [rank 0 half][rank 1 half]
  uint16_t n_slabs;
  uint16_t ksteps;      // k-steps (of 16) per slab
  uint8_t kind;         // K_FWD (N=256, 3-pass), K_RGB (N=128, 1-pass), K_BWD (N=256), K_BW4 (N=256 + 64 extra), K_BW0 (N=64)
  uint8_t a_src;        // A_ACT, A_PE, A_PE_ACT, A_ACT_LO
  uint8_t pad[2];
};


struct TcParams {
  TcOp ops[TC_NUM_OPS];     // the tile's GEMM schedule (kernel parameters live in the constant bank; no global state)
  const uint8_t* wpack;    // packed fp16 weights
  const float* bias;       // [7][256]; row 0 = per-frame folded bias (code + pose feature)
  const float* b_rgb1;     // [128]
  const float* w_rgb2;     // [3][128]
  const float* w_dens;     // [256]
  const uint32_t* seed_h2; // [128] packed half2 pairs of w_dens / seed_scale
  float b_rgb2[3];
  float b_dens;
  float seed_scale;        // out_g = accumulated gradient * seed_scale (undoes the seed normalisation and the per-layer power-of-two scales)
  float stash_scale;       // layer 4's PE-gradient (parked at op 10) carries fewer per-layer scales than layer 0's: multiply it by this
  const float4* active;
  const unsigned long long* n_active_ptr;
  int64_t n_active_host;
  float4* out_a;
  float4* out_g;
  int density_only;
  int rgb3;                // rgb head's 256 -> 128 layer with the 3-pass split (chosen by dsnerf_set_weights, see TcWeights::stage)
  long long* timing;       // debug: clock64 stamps of CTA 0 / first tile, NULL in production
  int debug_noload;        // debug: skip the weight stream (garbage results) to measure its cost
  int debug_passes;        // debug: 0 = normal; 1 / 2 = issue only that many MMAs per forward k-step (garbage results)
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// cta_group::2: arrives on the barrier at this CTA-relative offset in both CTAs of the pair once all prior MMAs are done
__device__ __forceinline__ void tc_commit2(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(mbar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]   (single CTA, used by light_tc.cuh)
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem] on the CTA pair: M = 256 (128 rows per CTA), B rows split between the two CTAs
__device__ __forceinline__ void tc_mma_ss2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// Remote arrive without a cluster-scope release: that form compiles to MEMBAR.ALL.GPU per arrival (and a cluster-scope
// acquire on the waiting side to CCTL.IVALL, an L1 flush), which doubled the epilogue time.  What is published here is this
// CTA's own shared memory, already made visible to the async proxy by fence.proxy.async, and is read by this SM's tensor core.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t a, uint32_t parity) { mbar_wait(a, parity); }  // barrier the peer CTA arrives on
// K-major, no swizzle: core matrix = 8 rows x 16 B contiguous; SBO = 128 B between 8-row groups,
// LBO = byte distance between the two 8-wide K chunks of one k-step.
#define DSN_R16(r) \
  "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), \
  "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
#define DSN_RW16(r) \
  "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), \
  "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
#define DSN_R32(r) \
  "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), \
  "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),  \
  "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),  \
  "=r"(r[31])
#define DSN_RW32(r) \
  "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), \
  "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),  \
  "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),  \
  "+r"(r[31])
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : DSN_R32(r) : "r"(taddr));
}
// wait for outstanding tcgen05.ld; the registers are in/out operands so the compiler cannot hoist a read above the wait
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[32]) { asm volatile("tcgen05.wait::ld.sync.aligned;" : DSN_RW32(r)::"memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  tmem_ld32_nowait(taddr, r);
  tmem_wait_ld(r);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : DSN_R16(r) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[16]) { asm volatile("tcgen05.wait::ld.sync.aligned;" : DSN_RW16(r)::"memory"); }
// one lane of a fully converged warp (the warp stays converged, so descriptors live in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_WARPS * 32) : "memory"); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ __half2 as_h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }

// ReLU bits of the 7 forward layers for this thread's 64 columns (2 words per layer), kept in registers.
// Bit layout inside a word: pair j (columns 2j, 2j+1 of the thread's 16) of quarter q4 -> bits p and 16+p, p = j + 8*(q4&1).
struct ReluBits {
  uint32_t w[7][2];
  __device__ __forceinline__ void put(int layer, uint32_t v0, uint32_t v1) {
#pragma unroll
    for (int l = 0; l < 7; ++l)
      if (l == layer) { w[l][0] = v0; w[l][1] = v1; }
  }
  __device__ __forceinline__ void get(int layer, uint32_t& v0, uint32_t& v1) const {
    v0 = 0; v1 = 0;
#pragma unroll
    for (int l = 0; l < 7; ++l)
      if (l == layer) { v0 = w[l][0]; v1 = w[l][1]; }
  }
};
// 0xffff in each half whose ReLU bit is set (P = bit position of the pair, compile-time)
template <int P>
__device__ __forceinline__ uint32_t relu_mask2(uint32_t word) {
  if (P == 15) return ((word >> 15) & 0x00010001u) * 0xFFFFu;  // bit 15 / 31 would read as -0 == 0 in a half compare
  const uint32_t t = word & ((1u << P) | (1u << (16 + P)));
  return __hne2_mask(as_h2(t), as_h2(0u));
}

// Issue all MMAs of one weight slab (KSTEPS k-steps) and release its ring slot.  Shapes are template
// parameters so that every descriptor is "slab base + immediate": the issuing lane spends a couple of
// uniform-datapath adds per tcgen05.mma instead of rebuilding 64-bit descriptors.
//   a_word / a_lo_word / b_word: low 32 bits of the A-hi / A-lo / B-hi shared-memory descriptors at the slab's first k-step
template <int ROWS, int KSTEPS, int NPASS, int NMMA, bool EXTRA>
__device__ __forceinline__ void issue_slab(uint32_t d_main, uint32_t d_extra, uint32_t a_word, uint32_t a_lo_word, uint32_t b_word,
                                           uint32_t first_acc, uint32_t empty_bar, bool do_commit = true) {  // ROWS = B rows held by ONE CTA
  constexpr uint32_t DHI = (128u >> 4) | (1u << 14);             // descriptor bits 32..63: SBO = 128 B, version 1
  constexpr uint32_t A_STEP = (2 * A_CHUNK) >> 4;                // one k-step along K in the A operand
  constexpr uint32_t B_STEP = (2 * ROWS * 16) >> 4;              // one k-step in the slab
  constexpr uint32_t B_LO = (KSTEPS * 2 * ROWS * 16) >> 4;       // hi part -> lo part of the slab
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NMMA >> 3) << 17) | ((256u >> 4) << 24);   // M = 256 over the CTA pair
  constexpr uint32_t IDESC_X = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((256u >> 4) << 24);
  if (elect_one()) {
#pragma unroll
    for (int j = 0; j < KSTEPS; ++j) {
      const uint64_t da = ((uint64_t)DHI << 32) | (a_word + j * A_STEP);
      const uint64_t db = ((uint64_t)DHI << 32) | (b_word + j * B_STEP);
      tc_mma_ss2(d_main, da, db, IDESC, j == 0 ? first_acc : 1u);
      if (NPASS >= 2) {
        const uint64_t dbl = ((uint64_t)DHI << 32) | (b_word + j * B_STEP + B_LO);
        tc_mma_ss2(d_main, da, dbl, IDESC, 1u);
      }
      if (NPASS >= 3) {
        const uint64_t dal = ((uint64_t)DHI << 32) | (a_lo_word + j * A_STEP);
        tc_mma_ss2(d_main, dal, db, IDESC, 1u);
      }
      if (EXTRA) {
        const uint64_t dbx = ((uint64_t)DHI << 32) | (b_word + j * B_STEP + ((128 * 16) >> 4));  // this CTA's 32 extra rows follow its 128 main rows
        tc_mma_ss2(d_extra, da, dbx, IDESC_X, j == 0 ? first_acc : 1u);
      }
    }
    if (do_commit) tc_commit2(empty_bar);  // frees the ring slot in both CTAs once these MMAs have read it
  }
  __syncwarp();
}

// State of the MMA-issuing warp (leader CTA) and the per-op slab loop.
struct MmaState {
  uint32_t stage, phase, a_phase, sbase, bar_full, bar_empty, bar_a;
  int waited, noload;
  // the producers' epilogues publish their 256 output columns in four quarters of 64 (= 4 k-steps of the A operand)
  __device__ __forceinline__ void need_quarters(int nq) {
    if (waited >= nq) return;
    while (waited < nq) { mbar_wait(bar_a + 8 * waited, a_phase); ++waited; }
    tc_fence_after();
  }
};
template <int ROWS, int KSTEPS, int NPASS, int NMMA, bool EXTRA>
__device__ __forceinline__ void run_op(MmaState& ms, int n_slabs, int a_src, uint32_t d_main, uint32_t d_extra) {
  constexpr uint32_t A_LBO = (A_CHUNK >> 4) << 16;          // LBO field of every A descriptor
  constexpr uint32_t B_LBO = ((ROWS * 16) >> 4) << 16;
  if (a_src == A_PE) ms.need_quarters(4);
  const uint32_t pe_steps = (a_src == A_PE || a_src == A_PE_ACT) ? 4u : 0u;  // leading k-steps served by the PE region
  const uint32_t a_hi0 = A_LBO | ((ms.sbase + (a_src == A_ACT_LO ? SM_A_LO : SM_A_HI)) >> 4), a_lo0 = A_LBO | ((ms.sbase + SM_A_LO) >> 4);
  const uint32_t p_hi0 = A_LBO | ((ms.sbase + SM_PE_HI) >> 4), p_lo0 = A_LBO | ((ms.sbase + SM_PE_LO) >> 4);
  uint32_t kk = 0;
  for (int s = 0; s < n_slabs; ++s, kk += KSTEPS) {
    const bool from_pe = kk < pe_steps;
    // the weight slab first (it landed long ago: the ring runs ahead), the descriptors next, and only then the wait that is
    // on the critical path -- the operand quarter from the epilogue -- so that nothing but the MMA issue follows it
    if (!ms.noload) mbar_wait(ms.bar_full + 8 * ms.stage, ms.phase);
    const uint32_t b_word = B_LBO | ((ms.sbase + SM_RING + ms.stage * TC_STAGE_BYTES) >> 4);
    const uint32_t ac = ((from_pe ? kk : kk - pe_steps) * 2 * A_CHUNK) >> 4;
    if (!from_pe) ms.need_quarters(min(4, (int)((kk - pe_steps + KSTEPS - 1) >> 2) + 1));
    tc_fence_after();
    issue_slab<ROWS, KSTEPS, NPASS, NMMA, EXTRA>(d_main, d_extra, (from_pe ? p_hi0 : a_hi0) + ac, (from_pe ? p_lo0 : a_lo0) + ac, b_word,
                                                 (uint32_t)(kk > 0), ms.bar_empty + 8 * ms.stage, !(ms.noload & 2));
    if (++ms.stage == TC_STAGES) { ms.stage = 0; ms.phase ^= 1; }
  }
}

// octaves of the positional encoding owned by column sub-block `sub` (of 4): [pe_k0(sub), pe_k0(sub+1))
__device__ __forceinline__ int pe_k0(int sub) { return sub == 0 ? 0 : (sub == 1 ? 2 : (sub == 2 ? 5 : (sub == 3 ? 8 : 10))); }
// first of the 32 consecutive PE-gradient columns a sub-block reads to cover its own columns
// (sub 0: 0..14, sub 1: 15..32, sub 2: 33..50, sub 3: 51..63)
__device__ __forceinline__ int pe_ld0(int sub) { return sub == 0 ? 0 : (sub == 1 ? 8 : 32); }

// sin / cos of 2^k * x for the positional encoding, ~30 instructions instead of sincosf's ~60 (the encoding is 30 sincos per
// point, ~4.5 k cycles of the whole SM per tile with sincosf).  y = x / 2pi is held as a two-float (yh, yl); 2^k * y is exact,
// its fraction is reduced to |r| <= 1/8 turn exactly, and sin / cos (2 pi r) are degree-9 / degree-8 polynomials in r with the
// low part of r folded in.  Max |error| 8e-8 against float64 over |x| <= 30, k = 0..9 (tests/test_host.py restates and checks
// it; CUDA's sincosf is <= 1.2e-7 there).  The reference's torch.sin / cos (model/dimension_kernel.py:27-33) carry ~6e-8.
__device__ __forceinline__ void pe_turns(float x, float& yh, float& yl) {
  const float chi = 0.15915493667125702f, clo = 6.4206382432985265e-09f;  // 1 / 2pi = chi + clo
  yh = __fmul_rn(x, chi);
  yl = __fmaf_rn(x, clo, __fmaf_rn(x, chi, -yh));
}
__device__ __forceinline__ void pe_sincos(float yh, float yl, int k, float& sn, float& cs) {
  const float sc = __int_as_float((127 + k) << 23);  // 2^k
  const float th = yh * sc, tl = yl * sc;             // exact
  const float fh = th - rintf(th);                    // exact, |fh| <= 1/2
  const float q = rintf(fh * 4.0f);                   // quarter turns, -2..2
  const float rh = __fmaf_rn(q, -0.25f, fh);          // exact, |rh| <= 1/8
  const float s = __fadd_rn(rh, tl);
  const float e = __fsub_rn(tl, __fsub_rn(s, rh));    // s + e = rh + tl
  const float u = s * s;
  float ps = 41.46822738647461f;
  ps = __fmaf_rn(ps, u, -76.69773864746094f);
  ps = __fmaf_rn(ps, u, 81.6052017211914f);
  ps = __fmaf_rn(ps, u, -41.34170150756836f);
  ps = __fmaf_rn(ps, u, -1.7484555e-07f);             // low part of 2 pi
  float pc = 59.41782760620117f;
  pc = __fmaf_rn(pc, u, -85.44869995117188f);
  pc = __fmaf_rn(pc, u, 64.93936920166016f);
  pc = __fmaf_rn(pc, u, -19.739208221435547f);
  const float s0 = __fmaf_rn(s, ps, s * 6.2831854820251465f);
  const float c0 = __fmaf_rn(pc, u, 1.0f);
  const float e2 = e * 6.2831854820251465f;
  const float s1 = __fmaf_rn(e2, c0, s0), c1 = __fmaf_rn(-e2, s0, c0);
  const int qi = (int)q & 3;
  const float a = (qi & 1) ? c1 : s1, b = (qi & 1) ? s1 : c1;  // (sin, cos) before the signs
  sn = (qi & 2) ? -a : a;
  cs = ((qi + 1) & 2) ? -b : b;
}

// ------------------------------------------------------------------------------------------ kernel
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) mlp_tc_kernel(TcParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + SM_BAR;              // [TC_STAGES]   half slab landed (leader: its own AND the peer's, 2 arrivals)
  const uint32_t bar_empty = bar_full + 8 * TC_STAGES;   // [TC_STAGES]   ring slot consumed (commit multicast to both CTAs)
  const uint32_t bar_acc = bar_empty + 8 * TC_STAGES;    //               accumulator complete (MMA -> epilogue, both CTAs)
  const uint32_t bar_a = bar_acc + 8;                    // [4]           (leader) both epilogues done with column quarter q
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * (2 * TC_STAGES + 6));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(bar_full + 8 * s, cta_rank == 0 ? 2 : 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_acc, 1);
    for (int q4 = 0; q4 < 4; ++q4) mbar_init(bar_a + 8 * q4, 2 * TC_EPI_WARPS);  // one elected lane per epilogue warp of both CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_EPI_WARPS) {  // same warp id in both CTAs of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 512; i += TC_THREADS)
    reinterpret_cast<float*>(smem + SM_RGBW)[i] = i < 128 ? P.b_rgb1[i] : P.w_rgb2[i - 128];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anyone arrives on them remotely
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (P.timing && blockIdx.x == 0 && threadIdx.x == 0) {  // SM clock inside the kernel: clock64 against the global timer
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    P.timing[120] = clock64();
    P.timing[121] = (long long)gt;
  }

  const int64_t n_active = P.n_active_ptr ? (int64_t)*P.n_active_ptr : P.n_active_host;
  const int64_t n_tiles = (n_active + TC_TILE - 1) / TC_TILE;
  // both CTAs of a pair run the leader's iteration count (the peer's last tile may be a dummy)
  const int64_t lead = blockIdx.x & ~1u;
  const int64_t n_iter = n_tiles > lead ? (n_tiles - lead + gridDim.x - 1) / gridDim.x : 0;
  const int n_ops = P.density_only ? 7 : TC_NUM_OPS;

  if (warp == TC_EPI_WARPS + 1) {
    // =============================== weight loader (one elected lane of a converged warp) =======
    uint32_t stage = 0, phase = 0;
    for (int64_t it = 0; it < n_iter; ++it) {
      for (int op = 0; op < n_ops; ++op) {
        const TcOp o = P.ops[op];
        const uint8_t* src = P.wpack + o.src_off + (size_t)cta_rank * o.slab_bytes;
        for (int s = 0; s < o.n_slabs; ++s) {
          if (P.debug_noload) continue;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bar_full + 8 * stage, o.slab_bytes);
            bulk_g2s(sbase + SM_RING + stage * TC_STAGE_BYTES, src + (size_t)s * 2 * o.slab_bytes, o.slab_bytes, bar_full + 8 * stage);
          }
          __syncwarp();
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS && cta_rank != 0) {
    // =============================== peer CTA: relay "my half of the slab has landed" to the leader ==========
    uint32_t stage = 0, phase = 0;
    const uint32_t leader_pfull = mapa_u32(bar_full, 0);
    for (int64_t it = 0; it < n_iter; ++it) {
      for (int op = 0; op < n_ops; ++op) {
        const int n_slabs = P.ops[op].n_slabs;
        for (int s = 0; s < n_slabs; ++s) {
          if (P.debug_noload) continue;
          mbar_wait(bar_full + 8 * stage, phase);
          if (lane == 0) mbar_arrive_cluster(leader_pfull + 8 * stage);
          __syncwarp();
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS) {
    // =============================== leader CTA: MMA issuer (one elected lane of a converged warp) ==========
    // Measured: one iteration of the slab loop costs the issuing warp ~450 cycles of dependent scalar work whatever it
    // issues, so a slab must carry more MMA time than that (2 forward k-steps = 6 MMAs = 768 cycles) and the loop body
    // is kept minimal: kind switch hoisted out of it, one "landed" barrier per slab, shapes as template parameters.
    MmaState ms;
    ms.stage = 0; ms.phase = 0; ms.a_phase = 0;
    ms.sbase = sbase; ms.bar_full = bar_full; ms.bar_empty = bar_empty; ms.bar_a = bar_a; ms.noload = P.debug_noload;
    for (int64_t it = 0; it < n_iter; ++it) {
      for (int op = 0; op < n_ops; ++op) {
        const TcOp o = P.ops[op];
        const bool mstamp = P.timing && blockIdx.x == 0 && it == 0 && lane == 0;
        const long long t_op0 = mstamp ? clock64() : 0;
        const uint32_t d_main = tmem + (uint32_t)(op & 1) * TM_ACC;
        const uint32_t d_extra = tmem + (uint32_t)((op & 1) ^ 1) * TM_ACC;
        // op 8 (bW6) reads the seed that layer 6's epilogue published together with h6 (consumed by op 7); with the 3-pass rgb
        // head the A-lo region holds h6's lo part until op 7 is done, and op 7's epilogue publishes the seed as a hand-off of its own
        ms.waited = (op == 8 && !P.rgb3) ? 4 : 0;
        switch (o.kind) {
          case K_FWD:
            // debug_passes (measurement only, wrong numerics): issue 2 or 1 of the 3 MMAs of a forward k-step = the time any
            // cheaper operand split could reach at best with this pipeline
            if (P.debug_passes == 2) run_op<128, 2, 2, 256, false>(ms, o.n_slabs, o.a_src, d_main, 0u);
            else if (P.debug_passes == 1) run_op<128, 2, 1, 256, false>(ms, o.n_slabs, o.a_src, d_main, 0u);
            else run_op<128, 2, 3, 256, false>(ms, o.n_slabs, o.a_src, d_main, 0u);
            break;
          case K_RGB: run_op<64, 8, 1, 128, false>(ms, o.n_slabs, o.a_src, d_main, 0u); break;
          case K_RGB3: run_op<64, 4, 3, 128, false>(ms, o.n_slabs, o.a_src, d_main, 0u); break;
          case K_BWD: run_op<128, 4, 1, 256, false>(ms, o.n_slabs, o.a_src, d_main, 0u); break;
          case K_BW4: run_op<160, 2, 1, 256, true>(ms, o.n_slabs, o.a_src, d_main, d_extra); break;
          default: run_op<32, 16, 1, 64, false>(ms, o.n_slabs, o.a_src, d_main, 0u); break;
        }
        ms.need_quarters(4);
        if (elect_one()) tc_commit2(bar_acc);  // accumulator (and the extra columns of K_BW4) complete, in both CTAs
        __syncwarp();
        if (op != 8 || P.rgb3) ms.a_phase ^= 1;
        if (mstamp) { P.timing[64 + 3 * op] = 0; P.timing[65 + 3 * op] = 0; P.timing[66 + 3 * op] = clock64() - t_op0; }
      }
    }
  } else {
    // =============================== epilogue warps ===========================================
    // warp w: TMEM lane quarter q = w % 4 (rows 32q..32q+31), column sub-block sub = w / 4: of every 64-column
    // accumulator quarter this thread handles columns [16*sub, 16*sub + 16) of its row.
    const int q = warp & 3, sub = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t row_off = (uint32_t)row * 16;
    uint32_t acc_phase = 0;
    ReluBits relu;
#pragma unroll
    for (int l = 0; l < 7; ++l) { relu.w[l][0] = 0; relu.w[l][1] = 0; }
    // publish quarter q4 of the A operand this warp has just written: writes -> async proxy, then one arrival per warp
    const uint32_t leader_bar_a = mapa_u32(bar_a, 0);
    auto publish = [&](int q4) {
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_bar_a + 8 * q4);
    };
    float4 pt_next = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((int64_t)blockIdx.x * TC_TILE + row < n_active) pt_next = P.active[(int64_t)blockIdx.x * TC_TILE + row];
    const int k_lo = pe_k0(sub), k_hi = pe_k0(sub + 1);
    float* stash = reinterpret_cast<float*>(smem + SM_STASH);  // [64 PE columns][128 rows] fp32, backward chain only
    float* xch = reinterpret_cast<float*>(smem + SM_XCH);      // [3][8][128]
    for (int64_t it = 0; it < n_iter; ++it) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int64_t base = tile * TC_TILE;
      const bool live = base + row < n_active;
      const float4 pt = pt_next;
      {  // fetch the next tile's point now: its DRAM latency hides behind this tile
        const int64_t nb = (tile + gridDim.x) * TC_TILE + row;
        pt_next = nb < n_active ? P.active[nb] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float xs[3] = {pt.x, pt.y, pt.z};
      // ---- positional encoding (model/dimension_kernel.py:5-35) = A operand of layer 0 and head of layer 4's.  Column c of
      // the 64-wide operand: c < 3: x_c; c = 3 + 6k + r: sin(2^k x_r) for r < 3, cos(2^k x_(r-3)) otherwise; c = 63: padding.
      // Column sub-block `sub` produces columns [16 sub, 16 sub + 16) = two 8-column chunks of its row, so that the operand is
      // written with four conflict-free 16-byte stores per thread (hi and lo), like every other A operand of the kernel;
      // per-element 2-byte stores into this layout are 4-way bank conflicted (47 M conflicts per launch in the round-1 profile).
      {
        float yh[3], yl[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) pe_turns(xs[c], yh[c], yl[c]);
        float v[16];
        float sn, cs;
#define DSN_SC(K, C) pe_sincos(yh[C], yl[C], K, sn, cs)
        if (sub == 0) {          // columns 0..15: x, y, z, octaves 0 and 1, sin(4 x)
          v[0] = xs[0]; v[1] = xs[1]; v[2] = xs[2];
#pragma unroll
          for (int c = 0; c < 3; ++c) { DSN_SC(0, c); v[3 + c] = sn; v[6 + c] = cs; }
#pragma unroll
          for (int c = 0; c < 3; ++c) { DSN_SC(1, c); v[9 + c] = sn; v[12 + c] = cs; }
          DSN_SC(2, 0); v[15] = sn;
        } else if (sub == 1) {   // columns 16..31: rest of octave 2, octave 3, octave 4 without cos(16 z)
          DSN_SC(2, 0); v[2] = cs;
          DSN_SC(2, 1); v[0] = sn; v[3] = cs;
          DSN_SC(2, 2); v[1] = sn; v[4] = cs;
#pragma unroll
          for (int c = 0; c < 3; ++c) { DSN_SC(3, c); v[5 + c] = sn; v[8 + c] = cs; }
          DSN_SC(4, 0); v[11] = sn; v[14] = cs;
          DSN_SC(4, 1); v[12] = sn; v[15] = cs;
          DSN_SC(4, 2); v[13] = sn;
        } else if (sub == 2) {   // columns 32..47: cos(16 z), octaves 5 and 6, the sines of octave 7
          DSN_SC(4, 2); v[0] = cs;
#pragma unroll
          for (int c = 0; c < 3; ++c) { DSN_SC(5, c); v[1 + c] = sn; v[4 + c] = cs; }
#pragma unroll
          for (int c = 0; c < 3; ++c) { DSN_SC(6, c); v[7 + c] = sn; v[10 + c] = cs; }
#pragma unroll
          for (int c = 0; c < 3; ++c) { DSN_SC(7, c); v[13 + c] = sn; }
        } else {                 // columns 48..63: the cosines of octave 7, octaves 8 and 9, padding
#pragma unroll
          for (int c = 0; c < 3; ++c) { DSN_SC(7, c); v[c] = cs; }
#pragma unroll
          for (int c = 0; c < 3; ++c) { DSN_SC(8, c); v[3 + c] = sn; v[6 + c] = cs; }
#pragma unroll
          for (int c = 0; c < 3; ++c) { DSN_SC(9, c); v[9 + c] = sn; v[12 + c] = cs; }
          v[15] = 0.f;
        }
#undef DSN_SC
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __half2 hh = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
          hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
          const float2 hf = __half22float2(hh);
          lo[j] = pack_h2(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
        }
        const uint32_t off = (uint32_t)(2 * sub) * A_CHUNK + row_off;
        *reinterpret_cast<uint4*>(smem + SM_PE_HI + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(smem + SM_PE_HI + off + A_CHUNK) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<uint4*>(smem + SM_PE_LO + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(smem + SM_PE_LO + off + A_CHUNK) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
      const bool stamp = P.timing && blockIdx.x == 0 && it == 0 && threadIdx.x == 0;
      if (stamp) P.timing[0] = clock64();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) mbar_arrive_cluster(leader_bar_a + 8 * q4);
      }

      float2 sig2 = make_float2(0.f, 0.f);
      float2 e0 = make_float2(0.f, 0.f), e1 = make_float2(0.f, 0.f), e2 = make_float2(0.f, 0.f);
      for (int op = 0; op < n_ops; ++op) {
        const uint32_t t_accb = t_lane + (uint32_t)(op & 1) * TM_ACC + sub * TC_CPT;
        if (op == 7) {
          // ---------- rgb head tail only.  Nothing is published: bW6 (op 8) already runs on the seed that layer 6's
          // epilogue wrote next to h6.
          mbar_wait(bar_acc, acc_phase);
          acc_phase ^= 1;
          tc_fence_after();
          if (stamp) { P.timing[1 + 4 * op] = clock64(); P.timing[2 + 4 * op] = clock64(); }
          if (P.rgb3) {
            // the rgb MMAs have consumed h6 (hi and lo): the seed of the backward chain G6 = (w_dens / scale) * relu'(a6) now
            // replaces h6's lo part (op 8's operand), rebuilt from the ReLU bits of layer 6
            uint32_t m0, m1;
            relu.get(6, m0, m1);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const uint4 sd0 = __ldg(reinterpret_cast<const uint4*>(P.seed_h2 + (q4 * 64 + sub * TC_CPT) / 2));
              const uint4 sd1 = __ldg(reinterpret_cast<const uint4*>(P.seed_h2 + (q4 * 64 + sub * TC_CPT) / 2 + 4));
              uint32_t g[8] = {sd0.x, sd0.y, sd0.z, sd0.w, sd1.x, sd1.y, sd1.z, sd1.w};
              const uint32_t mw = (q4 >> 1) ? m1 : m0;
              if ((q4 & 1) == 0) {
                g[0] &= relu_mask2<0>(mw); g[1] &= relu_mask2<1>(mw); g[2] &= relu_mask2<2>(mw); g[3] &= relu_mask2<3>(mw);
                g[4] &= relu_mask2<4>(mw); g[5] &= relu_mask2<5>(mw); g[6] &= relu_mask2<6>(mw); g[7] &= relu_mask2<7>(mw);
              } else {
                g[0] &= relu_mask2<8>(mw); g[1] &= relu_mask2<9>(mw); g[2] &= relu_mask2<10>(mw); g[3] &= relu_mask2<11>(mw);
                g[4] &= relu_mask2<12>(mw); g[5] &= relu_mask2<13>(mw); g[6] &= relu_mask2<14>(mw); g[7] &= relu_mask2<15>(mw);
              }
              const uint32_t off = (uint32_t)((q4 * 64 + sub * TC_CPT) / 8) * A_CHUNK + row_off;
              *reinterpret_cast<uint4*>(smem + SM_A_LO + off) = make_uint4(g[0], g[1], g[2], g[3]);
              *reinterpret_cast<uint4*>(smem + SM_A_LO + off + A_CHUNK) = make_uint4(g[4], g[5], g[6], g[7]);
              publish(q4);
            }
          }
          // rgb head tail: relu(acc[0:128] + b) -> Linear(128,3) partials (model/spacenet.py:75-80); 2 x 16 columns per thread
          const float* rgbw = reinterpret_cast<const float*>(smem + SM_RGBW);
          uint32_t v0[16], v1[16];
          tmem_ld16_nowait(t_accb, v0);
          tmem_ld16_nowait(t_accb + 64, v1);
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            const int col0 = hq * 64 + sub * TC_CPT;
            float4 b[4], w0[4], w1[4], w2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              b[i] = *(reinterpret_cast<const float4*>(rgbw + col0) + i);
              w0[i] = *(reinterpret_cast<const float4*>(rgbw + 128 + col0) + i);
              w1[i] = *(reinterpret_cast<const float4*>(rgbw + 256 + col0) + i);
              w2[i] = *(reinterpret_cast<const float4*>(rgbw + 384 + col0) + i);
            }
            if (hq == 0) tmem_wait_ld(v0); else tmem_wait_ld(v1);
            const uint32_t(&v)[16] = hq ? v1 : v0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float2 ra = __fadd2_rn(make_float2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), make_float2(b[i].x, b[i].y));
              float2 rb = __fadd2_rn(make_float2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), make_float2(b[i].z, b[i].w));
              ra.x = fmaxf(ra.x, 0.f); ra.y = fmaxf(ra.y, 0.f); rb.x = fmaxf(rb.x, 0.f); rb.y = fmaxf(rb.y, 0.f);
              e0 = __ffma2_rn(make_float2(w0[i].x, w0[i].y), ra, e0); e0 = __ffma2_rn(make_float2(w0[i].z, w0[i].w), rb, e0);
              e1 = __ffma2_rn(make_float2(w1[i].x, w1[i].y), ra, e1); e1 = __ffma2_rn(make_float2(w1[i].z, w1[i].w), rb, e1);
              e2 = __ffma2_rn(make_float2(w2[i].x, w2[i].y), ra, e2); e2 = __ffma2_rn(make_float2(w2[i].z, w2[i].w), rb, e2);
            }
          }
          tc_fence_before();
          if (stamp) { P.timing[3 + 4 * op] = clock64(); P.timing[4 + 4 * op] = clock64(); }
          continue;
        }
        mbar_wait(bar_acc, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        if (op == 14) break;  // the last op's accumulator is consumed by the tile-output stage below
        if (op == 10) {
          // layer 4 backward also produced d sigma / d PE (64 columns) in the idle accumulator, which the next op's MMAs
          // overwrite: park the columns this thread will need for the chain rule (own octaves) before publishing anything
          uint32_t v[32];
          const int c0 = pe_ld0(sub);
          tmem_ld32(t_lane + (uint32_t)((op & 1) ^ 1) * TM_ACC + c0, v);
          const int cb = sub == 0 ? 0 : 3 + 6 * k_lo, ce = sub == 3 ? 63 : 3 + 6 * k_hi;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i >= cb && c0 + i < ce) stash[(c0 + i) * TC_TILE + row] = __uint_as_float(v[i]);
        }
        if (stamp) P.timing[1 + 4 * op] = clock64();
        uint32_t va[16], vb[16];
        tmem_ld16_nowait(t_accb, va);
        if (op <= 6) {
          // ---------- forward layer: bias + ReLU, record ReLU bits, split to fp16 hi / lo = next A operand (in place).
          // Layer 6 feeds the single-pass rgb head (hi only) and the density head (fp32, here); instead of the lo part it
          // writes the seed of the backward chain G6 = (w_dens / scale) * relu'(a6) into the A-lo region (op 8's operand).
          const float* __restrict__ bias = P.bias + op * 256 + sub * TC_CPT;
          const float* __restrict__ wdp = P.w_dens + sub * TC_CPT;
          const bool last6 = op == 6;
          const bool seed6 = last6 && !P.rgb3;  // layer 6 writes the backward seed instead of its lo part (single-pass rgb head)
          uint32_t mw0 = 0, mw1 = 0;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t(&v)[16] = (q4 & 1) ? vb : va;
            float4 b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) b[i] = __ldg(reinterpret_cast<const float4*>(bias + q4 * 64) + i);
            uint4 sd0 = make_uint4(0u, 0u, 0u, 0u), sd1 = sd0;
            if (seed6) {
              sd0 = __ldg(reinterpret_cast<const uint4*>(P.seed_h2 + (q4 * 64 + sub * TC_CPT) / 2));
              sd1 = __ldg(reinterpret_cast<const uint4*>(P.seed_h2 + (q4 * 64 + sub * TC_CPT) / 2 + 4));
            }
            const uint32_t sd[8] = {sd0.x, sd0.y, sd0.z, sd0.w, sd1.x, sd1.y, sd1.z, sd1.w};
            tmem_wait_ld(v);
            if (q4 < 3) { if (q4 & 1) tmem_ld16_nowait(t_accb + (q4 + 1) * 64, va); else tmem_ld16_nowait(t_accb + (q4 + 1) * 64, vb); }
            uint32_t hi[8], lo[8];
            uint32_t mw = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = b[j >> 1];
              const float2 b2 = (j & 1) ? make_float2(bb.z, bb.w) : make_float2(bb.x, bb.y);
              float2 h = __fadd2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), b2);
              h.x = fmaxf(h.x, 0.f); h.y = fmaxf(h.y, 0.f);
              const __half2 hh = __floats2half2_rn(h.x, h.y);
              hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
              // ReLU bit = (hi > 0); an activation below the fp16 subnormal range is 0 for the forward pass and the gradient alike
              const uint32_t m2 = __hgt2_mask(hh, as_h2(0u));
              const int p = j + 8 * (q4 & 1);
              mw |= m2 & ((1u << p) | (1u << (16 + p)));
              if (last6) {
                const float2 wd = __ldg(reinterpret_cast<const float2*>(wdp + q4 * 64 + 2 * j));
                sig2 = __ffma2_rn(wd, h, sig2);
              }
              if (seed6) {
                lo[j] = sd[j] & m2;
              } else {
                const float2 hf = __half22float2(hh);
                const float2 l = __ffma2_rn(hf, make_float2(-1.f, -1.f), h);  // exact: x = hi + lo up to 2^-22 relative
                lo[j] = pack_h2(l.x, l.y);
              }
            }
            if (q4 >> 1) mw1 |= mw; else mw0 |= mw;
            if (op != n_ops - 1) {  // density-only mode ends at layer 6: nothing consumes the operand (and the exchange area aliases A-lo)
              const uint32_t off = (uint32_t)((q4 * 64 + sub * TC_CPT) / 8) * A_CHUNK + row_off;
              *reinterpret_cast<uint4*>(smem + SM_A_HI + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(smem + SM_A_HI + off + A_CHUNK) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
              *reinterpret_cast<uint4*>(smem + SM_A_LO + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              *reinterpret_cast<uint4*>(smem + SM_A_LO + off + A_CHUNK) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
              publish(q4);
            }
            if (stamp && q4 == 1) P.timing[2 + 4 * op] = clock64();
          }
          relu.put(op, mw0, mw1);
        } else {
          // ---------- backward layer: G_{l-1} = (G_l W_l) * relu'(a_{l-1}), fp16 single pass, hi only
          uint32_t m0, m1;
          relu.get(13 - op, m0, m1);  // op 8 -> layer 5 ... op 13 -> layer 0
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t(&v)[16] = (q4 & 1) ? vb : va;
            tmem_wait_ld(v);
            if (q4 < 3) { if (q4 & 1) tmem_ld16_nowait(t_accb + (q4 + 1) * 64, va); else tmem_ld16_nowait(t_accb + (q4 + 1) * 64, vb); }
            const uint32_t mw = (q4 >> 1) ? m1 : m0;
            uint32_t g[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = pack_h2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            if ((q4 & 1) == 0) {
              g[0] &= relu_mask2<0>(mw); g[1] &= relu_mask2<1>(mw); g[2] &= relu_mask2<2>(mw); g[3] &= relu_mask2<3>(mw);
              g[4] &= relu_mask2<4>(mw); g[5] &= relu_mask2<5>(mw); g[6] &= relu_mask2<6>(mw); g[7] &= relu_mask2<7>(mw);
            } else {
              g[0] &= relu_mask2<8>(mw); g[1] &= relu_mask2<9>(mw); g[2] &= relu_mask2<10>(mw); g[3] &= relu_mask2<11>(mw);
              g[4] &= relu_mask2<12>(mw); g[5] &= relu_mask2<13>(mw); g[6] &= relu_mask2<14>(mw); g[7] &= relu_mask2<15>(mw);
            }
            const uint32_t off = (uint32_t)((q4 * 64 + sub * TC_CPT) / 8) * A_CHUNK + row_off;
            *reinterpret_cast<uint4*>(smem + SM_A_HI + off) = make_uint4(g[0], g[1], g[2], g[3]);
            *reinterpret_cast<uint4*>(smem + SM_A_HI + off + A_CHUNK) = make_uint4(g[4], g[5], g[6], g[7]);
            publish(q4);
            if (stamp && q4 == 1) P.timing[2 + 4 * op] = clock64();
          }
        }
        if (stamp) { P.timing[3 + 4 * op] = clock64(); P.timing[4 + 4 * op] = clock64(); }
      }
      // ---------- tile outputs
      {
        float gx[3] = {0.f, 0.f, 0.f};
        if (!P.density_only) {
          // d sigma / d PE -> chain rule through the encoding for this thread's own octaves: layer-0 part = accumulator 0
          // (op 14, columns 0..63), layer-4 part = the stash written at op 10 (same thread, same columns)
          uint32_t g[32];
          tmem_ld32(t_lane + pe_ld0(sub), g);
          // sin / cos are still in the PE region: an octave's six columns lie in one or two 8-column chunks of the row, read
          // with 16-byte loads (hi and lo) -- per-element 2-byte loads of this layout are 4-way bank conflicted
          auto pe_chunk = [&](uint32_t region, int chunk) -> uint4 {
            return *reinterpret_cast<const uint4*>(smem + region + (uint32_t)chunk * A_CHUNK + row_off);
          };
          auto pe_pick = [](const uint4& hi, const uint4& lo, int e) -> float {  // element e of the chunk, hi + lo
            const int w = e >> 1;
            const uint32_t wh = w == 0 ? hi.x : (w == 1 ? hi.y : (w == 2 ? hi.z : hi.w));
            const uint32_t wl = w == 0 ? lo.x : (w == 1 ? lo.y : (w == 2 ? lo.z : lo.w));
            const uint32_t sh = (uint32_t)(e & 1) * 16u;
            return __half2float(__ushort_as_half((unsigned short)(wh >> sh))) + __half2float(__ushort_as_half((unsigned short)(wl >> sh)));
          };
          // every branch below is warp uniform (sub is); C0 = pe_ld0(sub) as a literal keeps g[] in registers
#define DSN_GPE(C, C0) fmaf(stash[(C) * TC_TILE + row], P.stash_scale, __uint_as_float(g[(C) - (C0)]))
#define DSN_OCTAVE(K, C0)                                                               \
  {                                                                                     \
    constexpr int CA = (3 + 6 * (K)) >> 3, CB = (8 + 6 * (K)) >> 3;                     \
    const uint4 ha = pe_chunk(SM_PE_HI, CA), la = pe_chunk(SM_PE_LO, CA);               \
    const uint4 hb = CB == CA ? ha : pe_chunk(SM_PE_HI, CB), lb = CB == CA ? la : pe_chunk(SM_PE_LO, CB); \
    _Pragma("unroll") for (int c = 0; c < 3; ++c) {                                     \
      const int cs_col = 3 + 6 * (K) + c, cc_col = 6 + 6 * (K) + c;                     \
      const float gs = DSN_GPE(3 + 6 * (K) + c, C0), gc = DSN_GPE(6 + 6 * (K) + c, C0); \
      const float sn = (cs_col >> 3) == CA ? pe_pick(ha, la, cs_col & 7) : pe_pick(hb, lb, cs_col & 7); \
      const float cs = (cc_col >> 3) == CA ? pe_pick(ha, la, cc_col & 7) : pe_pick(hb, lb, cc_col & 7); \
      gx[c] = fmaf((gs * cs - gc * sn), (float)(1 << (K)), gx[c]);                      \
    }                                                                                   \
  }
          if (sub == 0) {
            gx[0] = DSN_GPE(0, 0); gx[1] = DSN_GPE(1, 0); gx[2] = DSN_GPE(2, 0);
            DSN_OCTAVE(0, 0) DSN_OCTAVE(1, 0)
          } else if (sub == 1) {
            DSN_OCTAVE(2, 8) DSN_OCTAVE(3, 8) DSN_OCTAVE(4, 8)
          } else if (sub == 2) {
            DSN_OCTAVE(5, 32) DSN_OCTAVE(6, 32) DSN_OCTAVE(7, 32)
          } else {
            DSN_OCTAVE(8, 32) DSN_OCTAVE(9, 32)
          }
#undef DSN_OCTAVE
#undef DSN_GPE
        }
        // cross-thread sums (the four threads of a row live in warps q, q+4, q+8, q+12) through shared memory
        const float sigma_part = sig2.x + sig2.y;
        const float ee0 = e0.x + e0.y, ee1 = e1.x + e1.y, ee2 = e2.x + e2.y;
        if (sub != 0) {
          float* x = xch + (sub - 1) * 8 * TC_TILE + row;
          x[0] = sigma_part; x[TC_TILE] = ee0; x[2 * TC_TILE] = ee1; x[3 * TC_TILE] = ee2;
          x[4 * TC_TILE] = gx[0]; x[5 * TC_TILE] = gx[1]; x[6 * TC_TILE] = gx[2];
        }
        tc_fence_before();
        epi_bar();
        if (sub == 0 && live) {
          float s[7] = {sigma_part, ee0, ee1, ee2, gx[0], gx[1], gx[2]};
#pragma unroll
          for (int k = 0; k < 7; ++k) s[k] += xch[k * TC_TILE + row] + xch[(8 + k) * TC_TILE + row] + xch[(16 + k) * TC_TILE + row];
          const float sigma = s[0] + P.b_dens;
          if (P.density_only) {
            P.out_a[base + row] = make_float4(sigma, 0.f, 0.f, 0.f);
          } else {
            P.out_a[base + row] = make_float4(sigma, s[1] + P.b_rgb2[0], s[2] + P.b_rgb2[1], s[3] + P.b_rgb2[2]);
            P.out_g[base + row] = make_float4(s[4] * P.seed_scale, s[5] * P.seed_scale, s[6] * P.seed_scale, 0.f);
          }
        }
        if (stamp) P.timing[62] = clock64();
        // No second barrier: the exchange area, the stash and the PE region are rewritten by the next tile only after
        // MMAs that need an arrival from every epilogue warp, i.e. after every thread has left this block.
      }
    }
  }
  if (P.timing && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    P.timing[122] = clock64();
    P.timing[123] = (long long)gt;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the pair's MMAs / remote arrivals may still touch it
  if (warp == TC_EPI_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
struct TcWeights {
  void* d_pack = nullptr;
  void* d_pack2 = nullptr;   // the same weights in the op schedule of mlp_tc2.cuh (two tiles in flight)
  TcOp ops2[16];
  float* d_f32 = nullptr;  // biases 0..6 (7x256; row 0 per frame), b_rgb1 (128), w_rgb2 (384), w_dens (256), seed half2 pairs (128 words)
  float b_rgb2[3] = {0, 0, 0};
  float b_dens = 0.f, seed_scale = 1.f, stash_scale = 1.f;
  bool rgb3 = false;       // rgb head first layer packed for 3 passes (hi + lo)
  bool fp16_ok = true;     // false: some weight is outside fp16 range, the tensor-core path must not be used (dsnerf.cu routes to the fp32 kernel)
  int bwd_shift[7] = {0, 0, 0, 0, 0, 0, 0};  // per-layer power-of-two scale folded into the backward weights
  TcOp ops[TC_NUM_OPS];
  static constexpr int F32_BRGB1 = 7 * 256, F32_WRGB2 = F32_BRGB1 + 128, F32_WDENS = F32_WRGB2 + 384, F32_SEED = F32_WDENS + 256,
                       F32_TOTAL = F32_SEED + 128;
  float* bias0_slot() const { return d_f32; }

  void release() {
    if (d_pack) cudaFree(d_pack);
    if (d_pack2) cudaFree(d_pack2);
    if (d_f32) cudaFree(d_f32);
    d_pack = nullptr;
    d_pack2 = nullptr;
    d_f32 = nullptr;
  }

  // B[n][k] ((rows + extra) x K) packed as slabs of `ksteps` k-steps; every slab is split between the two CTAs of a pair:
  //   slab = [rank 0 half][rank 1 half], half = [hi: (2*ksteps chunks) x hrows x 8 halves][lo: same]
  // where a half holds rows/2 "main" rows (n = rank*rows/2 + i) followed by extra/2 "extra" rows
  // (n = rows + rank*extra/2 + i) -- exactly the image the cta_group::2 UMMA descriptors address in each CTA.
  static void pack_op(std::vector<__half>& blob, TcOp& op, int kind, int a_src, int rows, int extra, int K, int ksteps, bool with_lo,
                      const std::vector<float>& B) {
    const int n_slabs = K / (16 * ksteps);
    const int hmain = rows / 2, hextra = extra / 2, hrows = hmain + hextra;
    const size_t part = (size_t)ksteps * 2 * hrows * 8;  // halves per hi (or lo) part of one CTA's share
    const size_t half = part * (with_lo ? 2 : 1);
    while (blob.size() % 64) blob.push_back(__float2half_rn(0.f));  // 128-byte aligned slabs
    op.src_off = (uint32_t)(blob.size() * sizeof(__half));
    op.slab_bytes = (uint32_t)(half * sizeof(__half));
    op.n_slabs = (uint16_t)n_slabs;
    op.ksteps = (uint16_t)ksteps;
    op.kind = (uint8_t)kind;
    op.a_src = (uint8_t)a_src;
    op.pad[0] = op.pad[1] = 0;
    const size_t base = blob.size();
    blob.resize(base + 2 * half * n_slabs);
    for (int r = 0; r < 2; ++r)
      for (int i = 0; i < hrows; ++i) {
        const int n = i < hmain ? r * hmain + i : rows + r * hextra + (i - hmain);
        for (int k = 0; k < K; ++k) {
          const int s = k / (16 * ksteps), j = (k / 16) % ksteps, cc = (k % 16) / 8, e = k % 8;
          const size_t off = base + ((size_t)s * 2 + r) * half + ((size_t)(j * 2 + cc) * hrows + i) * 8 + e;
          const float w = B[(size_t)n * K + k];
          const __half hh = __float2half_rn(w);
          blob[off] = hh;
          if (with_lo) blob[off + part] = __float2half_rn(w - __half2float(hh));
        }
      }
  }

  int stage(const std::vector<float>& w0, const std::vector<float>& w1, const std::vector<float>& w2, const std::vector<float>& w3,
            const std::vector<float>& w4, const std::vector<float>& w5, const std::vector<float>& w6, const std::vector<float>& b1,
            const std::vector<float>& b2, const std::vector<float>& b3, const std::vector<float>& b4, const std::vector<float>& b5,
            const std::vector<float>& b6, const std::vector<float>& wd, float bd, const std::vector<float>& wr1,
            const std::vector<float>& br1, const std::vector<float>& wr2, const std::vector<float>& br2, bool rgb_three_pass) {
    rgb3 = rgb_three_pass;
    const std::vector<float>* W[7] = {&w0, &w1, &w2, &w3, &w4, &w5, &w6};
    std::vector<__half> blob;
    std::vector<float> B;
    int oi = 0;
    // forward layers 0..6: N = 256, 3-pass (hi + lo), two k-steps per slab (16 KB per CTA)
    for (int l = 0; l < 7; ++l) {
      const int in_dim = l == 0 ? 87 : (l == 4 ? 319 : 256);
      const int K = l == 0 ? 64 : (l == 4 ? 320 : 256);
      B.assign((size_t)256 * K, 0.f);
      for (int n = 0; n < 256; ++n) {
        if (l == 0) {
          for (int k = 0; k < 63; ++k) B[(size_t)n * K + k] = w0[(size_t)n * 87 + 8 + k];
        } else if (l == 4) {  // K order: [PE 0..62, pad | h 0..255] so that the PE k-steps are issued first
          for (int k = 0; k < 63; ++k) B[(size_t)n * K + k] = w4[(size_t)n * 319 + 256 + k];
          for (int k = 0; k < 256; ++k) B[(size_t)n * K + 64 + k] = w4[(size_t)n * 319 + k];
        } else {
          for (int k = 0; k < 256; ++k) B[(size_t)n * K + k] = (*W[l])[(size_t)n * in_dim + k];
        }
      }
      pack_op(blob, ops[oi++], K_FWD, l == 0 ? A_PE : (l == 4 ? A_PE_ACT : A_ACT), 256, 0, K, 2, true, B);
    }
    // rgb head first layer: 256 -> 128
    B.assign((size_t)128 * 256, 0.f);
    for (int n = 0; n < 128; ++n)
      for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = wr1[(size_t)n * 256 + k];
    if (rgb3) pack_op(blob, ops[oi++], K_RGB3, A_ACT, 128, 0, 256, 4, true, B);
    else pack_op(blob, ops[oi++], K_RGB, A_ACT, 128, 0, 256, 8, false, B);
    // fp16 range: the forward operands are fp16 hi/lo splits of the weights, so |w| must be below 65504; the backward chain
    // G_{l-1} = (G_l W_l) * relu' repacks G to fp16 at every layer, so each backward weight matrix carries a power-of-two scale
    // 2^shift[l] ~ 1 / (rms gain of the layer) that keeps |G| near the seed's magnitude whatever the checkpoint's weight norms are
    // (the normal is scale invariant; out_g is rescaled by seed_scale).  Scaling by powers of two is exact.
    fp16_ok = true;
    for (int l = 0; l < 7; ++l) {
      const int in_dim = l == 0 ? 87 : (l == 4 ? 319 : 256);
      double ss = 0.0, mx = 0.0;
      for (int n = 0; n < 256; ++n)
        for (int k = 0; k < in_dim; ++k) {
          if (l == 0 && (k < 8 || k >= 71)) continue;  // code / pose columns are folded into the bias
          const double v = (*W[l])[(size_t)n * in_dim + k];
          ss += v * v;
          mx = std::max(mx, fabs(v));
        }
      if (!(mx < 60000.0)) fp16_ok = false;
      const double cols = l == 0 ? 63 : in_dim;
      const double gain = sqrt(0.5 * ss / cols);        // rms of one backward output per unit rms input, half of the units active
      int sh = gain > 0.0 ? (int)lrint(-log2(gain)) : 0;
      sh = std::max(-12, std::min(12, sh));
      while (sh > -12 && mx * exp2((double)sh) > 30000.0) --sh;
      bwd_shift[l] = sh;
    }
    for (float v : wd) if (!(fabsf(v) < 3.0e38f)) fp16_ok = false;
    for (float v : wr1) if (!(fabsf(v) < 60000.f)) fp16_ok = false;
    auto p2 = [](int sh) { return (float)exp2((double)sh); };
    // backward through layers 6..1: B[n][k] = W[k][n] (n = input index, k = output index), 1-pass
    for (int l = 6; l >= 1; --l) {
      const float sc = p2(bwd_shift[l]);
      if (l == 4) {  // 256 hidden inputs + 63 PE inputs (+1 pad): rows 256..319 feed the extra N = 64 MMA
        B.assign((size_t)320 * 256, 0.f);
        for (int n = 0; n < 319; ++n)
          for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = sc * w4[(size_t)k * 319 + n];
        pack_op(blob, ops[oi++], K_BW4, A_ACT, 256, 64, 256, 2, false, B);
      } else {
        B.assign((size_t)256 * 256, 0.f);
        for (int n = 0; n < 256; ++n)
          for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = sc * (*W[l])[(size_t)k * 256 + n];
        pack_op(blob, ops[oi++], K_BWD, l == 6 ? A_ACT_LO : A_ACT, 256, 0, 256, 4, false, B);
      }
    }
    // layer 0 backward, PE columns only (N = 64); added to the stashed layer-4 PE gradient in the tile's last stage
    B.assign((size_t)64 * 256, 0.f);
    for (int n = 0; n < 63; ++n)
      for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = p2(bwd_shift[0]) * w0[(size_t)k * 87 + 8 + n];
    pack_op(blob, ops[oi++], K_BW0, A_ACT, 64, 0, 256, 16, false, B);
    if (oi != TC_NUM_OPS) return (int)cudaErrorUnknown;
    for (int i = 0; i < TC_NUM_OPS; ++i)
      if (ops[i].slab_bytes > TC_STAGE_BYTES || (ops[i].slab_bytes & 31) || (ops[i].src_off & 15)) return (int)cudaErrorInvalidValue;
    // ---- schedule of mlp_tc2.cuh (16 ops): slabs of 16 KB per CTA wherever the shape allows (the weight ring's throughput is
    // slabs in flight per round trip, whatever their size); the extra
    // N = 64 product of layer 4's backward pass (d sigma / d PE) is an op of its own in front of the main product
    std::vector<__half> blob2;
    {
      int o2 = 0;
      for (int l = 0; l < 7; ++l) {
        const int K = l == 0 ? 64 : (l == 4 ? 320 : 256);
        B.assign((size_t)256 * K, 0.f);
        for (int n = 0; n < 256; ++n) {
          if (l == 0) {
            for (int k = 0; k < 63; ++k) B[(size_t)n * K + k] = w0[(size_t)n * 87 + 8 + k];
          } else if (l == 4) {
            for (int k = 0; k < 63; ++k) B[(size_t)n * K + k] = w4[(size_t)n * 319 + 256 + k];
            for (int k = 0; k < 256; ++k) B[(size_t)n * K + 64 + k] = w4[(size_t)n * 319 + k];
          } else {
            for (int k = 0; k < 256; ++k) B[(size_t)n * K + k] = (*W[l])[(size_t)n * 256 + k];
          }
        }
        pack_op(blob2, ops2[o2++], K_FWD, 0, 256, 0, K, 2, true, B);
      }
      B.assign((size_t)128 * 256, 0.f);
      for (int n = 0; n < 128; ++n)
        for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = wr1[(size_t)n * 256 + k];
      pack_op(blob2, ops2[o2++], rgb3 ? K_RGB3 : K_RGB, 0, 128, 0, 256, rgb3 ? 4 : 8, rgb3, B);
      for (int l = 6; l >= 1; --l) {
        const float sc = p2(bwd_shift[l]);
        const int in_dim = l == 4 ? 319 : 256;
        if (l == 4) {  // d sigma / d PE through layer 4: rows = the 63 PE inputs
          B.assign((size_t)64 * 256, 0.f);
          for (int n = 0; n < 63; ++n)
            for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = sc * w4[(size_t)k * 319 + 256 + n];
          pack_op(blob2, ops2[o2++], K_BW0, 0, 64, 0, 256, 16, false, B);
        }
        B.assign((size_t)256 * 256, 0.f);
        for (int n = 0; n < 256; ++n)
          for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = sc * (*W[l])[(size_t)k * in_dim + n];
        pack_op(blob2, ops2[o2++], K_BWD, 0, 256, 0, 256, 4, false, B);
      }
      B.assign((size_t)64 * 256, 0.f);
      for (int n = 0; n < 63; ++n)
        for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = p2(bwd_shift[0]) * w0[(size_t)k * 87 + 8 + n];
      pack_op(blob2, ops2[o2++], K_BW0, 0, 64, 0, 256, 16, false, B);
      if (o2 != 16) return (int)cudaErrorUnknown;
      for (int i = 0; i < 16; ++i)
        if (ops2[i].slab_bytes > TC_STAGE_BYTES || (ops2[i].slab_bytes & 31) || (ops2[i].src_off & 15)) return (int)cudaErrorInvalidValue;
    }
    release();
    cudaError_t e = cudaMalloc(&d_pack, blob.size() * sizeof(__half));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(d_pack, blob.data(), blob.size() * sizeof(__half), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
    e = cudaMalloc(&d_pack2, blob2.size() * sizeof(__half));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(d_pack2, blob2.data(), blob2.size() * sizeof(__half), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
    std::vector<float> f(F32_TOTAL, 0.f);
    const std::vector<float>* bs[6] = {&b1, &b2, &b3, &b4, &b5, &b6};
    for (int l = 0; l < 6; ++l)
      for (int i = 0; i < 256; ++i) f[(l + 1) * 256 + i] = (*bs[l])[i];
    for (int i = 0; i < 128; ++i) f[F32_BRGB1 + i] = br1[i];
    for (int i = 0; i < 384; ++i) f[F32_WRGB2 + i] = wr2[i];
    float mx = 0.f;
    for (int i = 0; i < 256; ++i) mx = fmaxf(mx, fabsf(wd[i]));
    const float seed_norm = mx > 0.f ? mx : 1.f;
    for (int i = 0; i < 256; ++i) f[F32_WDENS + i] = wd[i];
    for (int i = 0; i < 128; ++i) {
      __half2 h = __floats2half2_rn(wd[2 * i] / seed_norm, wd[2 * i + 1] / seed_norm);
      memcpy(&f[F32_SEED + i], &h, 4);
    }
    {
      int all = 0, low = 0;  // layer 0's PE gradient went through every scale, layer 4's (stash) only through layers 6..4
      for (int l = 0; l < 7; ++l) all += bwd_shift[l];
      for (int l = 0; l < 4; ++l) low += bwd_shift[l];
      seed_scale = seed_norm * (float)exp2((double)-all);
      stash_scale = (float)exp2((double)low);
    }
    e = cudaMalloc(reinterpret_cast<void**>(&d_f32), f.size() * sizeof(float));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(d_f32, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
    for (int i = 0; i < 3; ++i) b_rgb2[i] = br2[i];
    b_dens = bd;
    return (int)cudaSuccess;
  }
};

inline void tc_configure() { cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM); }

inline int tc_launch(TcWeights& w, long long* timing, int debug_noload, int debug_passes, const float4* active, const unsigned long long* n_active_ptr, int64_t n_active_host,
                     float4* out_a, float4* out_g, int density_only, int sm_count, cudaStream_t st) {
  TcParams p{};
  for (int i = 0; i < TC_NUM_OPS; ++i) p.ops[i] = w.ops[i];
  p.wpack = reinterpret_cast<const uint8_t*>(w.d_pack);
  p.bias = w.d_f32;
  p.b_rgb1 = w.d_f32 + TcWeights::F32_BRGB1;
  p.w_rgb2 = w.d_f32 + TcWeights::F32_WRGB2;
  p.w_dens = w.d_f32 + TcWeights::F32_WDENS;
  p.seed_h2 = reinterpret_cast<const uint32_t*>(w.d_f32 + TcWeights::F32_SEED);
  for (int i = 0; i < 3; ++i) p.b_rgb2[i] = w.b_rgb2[i];
  p.b_dens = w.b_dens;
  p.seed_scale = w.seed_scale;
  p.stash_scale = w.stash_scale;
  p.active = active;
  p.n_active_ptr = n_active_ptr;
  p.n_active_host = n_active_host;
  p.out_a = out_a;
  p.out_g = out_g;
  p.density_only = density_only;
  p.rgb3 = w.rgb3 ? 1 : 0;
  p.timing = timing;
  p.debug_noload = debug_noload;
  p.debug_passes = debug_passes;
  mlp_tc_kernel<<<sm_count & ~1, TC_THREADS, TC_SMEM, st>>>(p);  // CTA pairs (__cluster_dims__(2,1,1))
  return (int)cudaGetLastError();
}

}  // namespace dsn

// SpaceNet forward + analytic density gradient on tcgen05, TWO tiles in flight per CTA (sm_100a).
//
// Same arithmetic as mlp_tc.cuh (model/spacenet.py:93-148 and :301-311: 3-pass fp16 hi/lo forward, single-pass fp16
// backward-data chain, fp32 heads) and the same CTA-pair / cta_group::2 / weight-stream design, but a different pipeline.
//
// Why.  mlp_tc.cuh runs ONE 128-point tile per CTA: the epilogue of layer l, the hand-off to the MMA warp, the MMAs of
// layer l + 1 and the hand-off back are a serial chain (profiles/r02/r02_pass_count_experiment.txt: a forward layer takes
// 8.5 k cycles for 6.1 k cycles of MMAs, a backward layer 4.1 k for 2.0 k; tensor pipe busy 59 %).  Only a second tile whose
// MMAs run under the first tile's epilogue removes that chain, and two tiles with a private, in-place A operand each
// (2 x 128 KB) do not fit in shared memory.
//
// How it fits here.
//   * The A operands of BOTH tiles live in one FIFO ring of ten 16 KB slots (a "unit" = 64 operand columns of a tile = one
//     hi slot [8 chunks][128 rows][8 halves], plus one lo slot for the 3-pass forward operands).  The epilogue warps produce
//     units in exactly the order in which the MMA warp consumes them (tile X layer l, tile Y layer l, X l+1, Y l+1, ...).
//     While the MMAs of (X, l+1) drain X's four units, the epilogue of (Y, l) refills the slots behind them: 160 KB hold
//     what would need 256 KB as private buffers.  A slot is handed back through a per-slot "free" barrier (tcgen05.commit
//     after the last MMA that read it) -- only where it is needed: a slot last read by a backward op is rewritten by a
//     phase that has already seen a later accumulator.  There is no per-unit "full" barrier: the accumulator release at the
//     end of a phase tells the MMA warp that all units of the phase are written and fenced (the encoding unit of op 0,
//     written a phase earlier, has a barrier of its own).
//   * Each tile owns ONE 256-column TMEM accumulator, used in place: the epilogue of (X, l) has read it completely before
//     the MMAs of (X, l+1) overwrite it ("acc free" barrier), and the MMAs of the other tile fill that time.
//   * Weights: forward layers stream their 16 KB half-slabs once per tile through a 4-stage ring; an op of the backward half
//     fits the ring, so its slabs are loaded once per tile PAIR (read for tile slot 0 without releasing, again for slot 1).
//   * The positional encoding is not kept: it is a unit like any other, produced when layer 0 and layer 4 need it
//     (recomputed; ~30 instructions per sincos), and the chain rule at the end of the gradient recomputes sin / cos with the
//     fast intrinsics (the gradient chain is fp16).
//   * d sigma / d PE of layer 4 (64 extra output columns of the backward pass through layer 4) has no spare TMEM columns:
//     it is an op of its own (N = 64) that reads the same four units as the main product WITHOUT releasing them (both tiles'
//     units fit: 8 of the 10 slots), and its chain rule is applied at once, so only three partial sums per thread survive.
//   * ReLU bits (14 words per thread and tile) go through a global scratch area (L2 resident, read back one backward layer
//     ahead) instead of registers.
//   * sigma and the rgb head's output leave the kernel at the rgb op, the gradient at the last op (partial sums of the four
//     threads of a row through TMEM: the accumulator columns that the N = 128 / N = 64 products of those ops leave unused;
//     the four threads of a row sit in warps q, q+4, q+8, q+12, which all reach TMEM lane quarter q).
//
// The kernel runs at the power cap: instruction count shows up as clock.  Hence wait loops without bookkeeping, nanosleep
// back-off where the waiter is ahead of the tensor pipe, layer 6 as its own instantiation instead of predicated code,
// and the MMA warp's loop reduced to one wait per slab (DESIGN.md 4b has the measurements).
//
// Op schedule of a tile (16 ops): 0-6 forward layers, 7 rgb head (its epilogue also writes the backward seed units),
// 8 bW6, 9 bW5, 10 d sigma / d PE through layer 4 (N = 64, no release), 11 bW4, 12 bW3, 13 bW2, 14 bW1, 15 bW0 (N = 64).
#pragma once
#include <type_traits>

#include "mlp_tc.cuh"

namespace dsn {

constexpr int T2_SLOTS = 10;
constexpr uint32_t T2_SLOT = 16384;
constexpr int T2_WST = 4;                                        // weight-ring stages (16 KB each)
constexpr uint32_t S2_A = 0;
constexpr uint32_t S2_RING = T2_SLOTS * T2_SLOT;                 // 163840
constexpr uint32_t S2_RGBW = S2_RING + T2_WST * TC_STAGE_BYTES;  // 229376: b_rgb1 [128], w_rgb2 [3][128]
constexpr uint32_t S2_BAR = S2_RGBW + 2048;                      // 231424
constexpr uint32_t T2_SMEM = S2_BAR + 272;                       // 231696 (32 barriers + the TMEM base slot)
constexpr uint32_t TM_XCH = 128;                                 // exchange columns inside a tile's accumulator (free at the rgb op and the last op)
constexpr int T2_NUM_OPS = 16;
#ifndef T2_SHARE_SLABS
#define T2_SHARE_SLABS 1   // ops 7..15 (at most four 16 KB slabs = the whole weight ring): both tiles use one copy of each slab
#endif
#ifndef T2_TIMING
#define T2_TIMING 0   // 1: per-phase clock64 stamps of CTA 0 (tests/tc2_timing.py needs a build with -DT2_TIMING=1)
#endif
constexpr int T2_STAMP_IT = 3;                                   // debug stamps are taken on this tile pair of CTA 0 (steady state)
constexpr uint32_t T2_KSTEP = 2 * A_CHUNK;                       // bytes per k-step inside a slot

struct Tc2Params {
  TcOp ops[T2_NUM_OPS];
  const uint8_t* wpack;
  const float* bias;        // [7][256]
  const float* b_rgb1;
  const float* w_rgb2;
  const float* w_dens;
  const uint32_t* seed_h2;
  float b_rgb2[3];
  float b_dens;
  float seed_scale;
  float stash_scale;
  const float4* active;
  const unsigned long long* n_active_ptr;
  int64_t n_active_host;
  float4* out_a;
  float4* out_g;
  uint32_t* relu_scratch;   // [grid][2 tiles][7 layers][2 words][512 threads]
  int rgb3;
  long long* timing;        // debug stamps of CTA 0 (NULL in production)
  unsigned int* dbg;        // watchdog record (NULL: spin for ever)
  int debug_noload;         // debug: skip the weight stream (garbage results) to measure its cost
};

// Barrier waits.  With a watchdog record (DSNERF_TC_WATCHDOG, bring-up only) a protocol error becomes a trap with a record
// instead of a hung GPU; the production loops carry no bookkeeping (every extra instruction of a wait loop is executed tens of
// millions of times per launch, and the kernel runs at the power cap).
__device__ __noinline__ void mbar_wait_watchdog(uint32_t a, uint32_t parity, unsigned int* dbg, uint32_t id) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 24)) {
      if (atomicCAS(dbg, 0u, 1u) == 0u) { dbg[1] = id; dbg[2] = blockIdx.x; dbg[3] = threadIdx.x; dbg[4] = parity; __threadfence(); }
      __trap();
    }
  } while (!done);
}
__device__ __forceinline__ void mbar_wait2(uint32_t a, uint32_t parity, unsigned int* dbg, uint32_t id) {
  if (dbg) { mbar_wait_watchdog(a, parity, dbg, id); return; }
  mbar_wait(a, parity);
}

#ifndef T2_WAIT_HINT
#define T2_WAIT_HINT 1000   // ns; 0 = plain polling
#endif
// epilogue-side wait: with a suspend-time hint the hardware parks the warp until the phase completes instead of re-issuing
// try_wait (every poll is a shared-memory access and an issue slot taken from the warps that work)
__device__ __forceinline__ void mbar_wait2h(uint32_t a, uint32_t parity, unsigned int* dbg, uint32_t id) {
  if (dbg) { mbar_wait_watchdog(a, parity, dbg, id); return; }
#if T2_WAIT_HINT > 0
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(a), "r"(parity), "r"((uint32_t)T2_WAIT_HINT) : "memory");
  } while (!done);
#else
  mbar_wait(a, parity);
#endif
}

// wait for a ring slot: the epilogue that waits here is ahead of the tensor pipe, so it can afford a coarse wake-up; every poll not
// issued is an instruction (and a shared-memory access) less on a kernel that runs at the power cap
#ifndef T2_FREE_SLEEP
#define T2_FREE_SLEEP 500   // ns between polls of a slot's "free" barrier (measured: 200 / 500 / 1000 alike, -1 % kernel time); 0 = same wait as everywhere else
#endif
__device__ __forceinline__ void mbar_wait_slot(uint32_t a, uint32_t parity, unsigned int* dbg, uint32_t id) {
#if T2_FREE_SLEEP > 0
  if (dbg) { mbar_wait_watchdog(a, parity, dbg, id); return; }
  uint32_t done;
  for (;;) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(T2_FREE_SLEEP);
  }
#else
  mbar_wait2h(a, parity, dbg, id);
#endif
}

// MMAs of (part of) one unit out of one weight slab; optional release of the unit's slots
template <int ROWS, int KSTEPS, int NPASS, int NMMA>
__device__ __forceinline__ void issue_unit(uint32_t d, uint32_t a_word, uint32_t a_lo_word, uint32_t b_word, uint32_t first_acc,
                                           uint32_t wempty_bar, uint32_t free0, uint32_t free1) {
  constexpr uint32_t DHI = (128u >> 4) | (1u << 14);
  constexpr uint32_t A_STEP = T2_KSTEP >> 4;
  constexpr uint32_t B_STEP = (2 * ROWS * 16) >> 4;
  constexpr uint32_t B_LO = (KSTEPS * 2 * ROWS * 16) >> 4;
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NMMA >> 3) << 17) | ((256u >> 4) << 24);
  if (elect_one()) {
#pragma unroll
    for (int j = 0; j < KSTEPS; ++j) {
      const uint64_t da = ((uint64_t)DHI << 32) | (a_word + j * A_STEP);
      const uint64_t db = ((uint64_t)DHI << 32) | (b_word + j * B_STEP);
      tc_mma_ss2(d, da, db, IDESC, j == 0 ? first_acc : 1u);
      if (NPASS >= 2) {
        const uint64_t dbl = ((uint64_t)DHI << 32) | (b_word + j * B_STEP + B_LO);
        tc_mma_ss2(d, da, dbl, IDESC, 1u);
      }
      if (NPASS >= 3) {
        const uint64_t dal = ((uint64_t)DHI << 32) | (a_lo_word + j * A_STEP);
        tc_mma_ss2(d, dal, db, IDESC, 1u);
      }
    }
    if (wempty_bar) tc_commit2(wempty_bar);
    if (free0) tc_commit2(free0);
    if (free1) tc_commit2(free1);
  }
  __syncwarp();
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16x(uint32_t taddr, uint32_t (&r)[16]) {
  tmem_ld16_nowait(taddr, r);
  tmem_wait_ld(r);
}

__device__ __forceinline__ void sincos_turns_fast(float yh, float yl, int k, float& sn, float& cs) {
  const float sc = __int_as_float((127 + k) << 23);
  const float th = yh * sc;
  const float r = (th - rintf(th)) + yl * sc;   // |r| <= 1/2 turn (+ a rounding)
  const float a = r * 6.2831854820251465f;
  sn = __sinf(a);
  cs = __cosf(a);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) mlp_tc2_kernel(Tc2Params P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_wfull = sbase + S2_BAR;               // [4]  half slab landed (leader: own + peer, 2 arrivals)
  const uint32_t bar_wempty = bar_wfull + 8 * T2_WST;      // [4]  weight slot consumed (commit, both CTAs)
  const uint32_t bar_pefull = bar_wempty + 8 * T2_WST;     // [2 of 10 entries] (leader) encoding unit of tile slot s (operand of op 0) written by all epilogue warps of the pair
  const uint32_t bar_afree = bar_pefull + 8 * T2_SLOTS;     // [10] slot read by its last MMA (commit, both CTAs)
  const uint32_t bar_accfull = bar_afree + 8 * T2_SLOTS;   // [2]  accumulator of tile slot s complete (commit, both CTAs)
  const uint32_t bar_accfree = bar_accfull + 16;           // [2]  (leader) accumulator of tile slot s read by every epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S2_BAR + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  unsigned int* const dbg = P.dbg;

  if (threadIdx.x == 0) {
    for (int s = 0; s < T2_WST; ++s) { mbar_init(bar_wfull + 8 * s, cta_rank == 0 ? 2 : 1); mbar_init(bar_wempty + 8 * s, 1); }
    for (int s = 0; s < T2_SLOTS; ++s) { mbar_init(bar_pefull + 8 * s, 2 * TC_EPI_WARPS); mbar_init(bar_afree + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_accfull + 8 * s, 1); mbar_init(bar_accfree + 8 * s, 2 * TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 512; i += TC_THREADS)
    reinterpret_cast<float*>(smem + S2_RGBW)[i] = i < 128 ? P.b_rgb1[i] : P.w_rgb2[i - 128];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (P.timing && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    P.timing[120] = clock64();
    P.timing[121] = (long long)gt;
  }

  const int64_t n_active = P.n_active_ptr ? (int64_t)*P.n_active_ptr : P.n_active_host;
  const int64_t n_tiles = (n_active + TC_TILE - 1) / TC_TILE;
  // both CTAs of a pair run the leader's iteration count; an iteration = two tiles per CTA (tile slots 0 and 1)
  const int64_t lead = blockIdx.x & ~1u;
  const int64_t cnt = n_tiles > lead ? (n_tiles - lead + gridDim.x - 1) / gridDim.x : 0;
  const int64_t n_iter = (cnt + 1) >> 1;

  if (warp == TC_EPI_WARPS + 1) {
    // =============================== weight loader ===============================
    uint32_t stage = 0, phase = 0;
    for (int64_t it = 0; it < n_iter; ++it) {
      for (int op = 0; op < T2_NUM_OPS; ++op) {
        const TcOp o = P.ops[op];
        const uint8_t* src = P.wpack + o.src_off + (size_t)cta_rank * o.slab_bytes;
        // forward layers stream their slabs once per tile; an op of the backward half fits the ring (<= 4 slabs), so its slabs are
        // loaded once per tile PAIR: the MMA warp reads them for tile slot 0 without releasing and again for tile slot 1
        const int passes = (T2_SHARE_SLABS && op >= 7) ? 1 : 2;
        for (int s2 = 0; s2 < passes; ++s2) {
          for (int s = 0; s < o.n_slabs; ++s) {
            if (P.debug_noload) continue;
            mbar_wait2(bar_wempty + 8 * stage, phase ^ 1, dbg, 0x100 + stage);
            if (elect_one()) {
              mbar_expect_tx(bar_wfull + 8 * stage, o.slab_bytes);
              bulk_g2s(sbase + S2_RING + stage * TC_STAGE_BYTES, src + (size_t)s * 2 * o.slab_bytes, o.slab_bytes, bar_wfull + 8 * stage);
            }
            __syncwarp();
            if (++stage == T2_WST) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS && cta_rank != 0) {
    // =============================== peer CTA: relay "my half of the slab has landed" ===============================
    uint32_t stage = 0, phase = 0;
    const uint32_t leader_wfull = mapa_u32(bar_wfull, 0);
    for (int64_t it = 0; it < n_iter; ++it) {
      for (int op = 0; op < T2_NUM_OPS; ++op) {
        const int n_slabs = ((T2_SHARE_SLABS && op >= 7) ? 1 : 2) * P.ops[op].n_slabs;
        for (int s = 0; s < n_slabs; ++s) {
          if (P.debug_noload) continue;
          mbar_wait2(bar_wfull + 8 * stage, phase, dbg, 0x200 + stage);
          if (lane == 0) mbar_arrive_cluster(leader_wfull + 8 * stage);
          __syncwarp();
          if (++stage == T2_WST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS) {
    // =============================== leader CTA: MMA issuer ===============================
    // The loop body is a chain of dependent scalar instructions issued from a scheduler that four busy epilogue warps share,
    // so it is kept minimal: the only waits are "accumulator free" per op (which also says that every unit the op reads has been
    // written and fenced: the epilogue releases the accumulator after its last unit), "encoding unit written" for op 0, and
    // "weight slab landed" per slab.
    constexpr uint32_t A_LBO = (A_CHUNK >> 4) << 16;
    const uint32_t a_base = A_LBO | ((sbase + S2_A) >> 4);
    const uint32_t w_base = (sbase + S2_RING) >> 4;
    uint32_t wstage = 0, wphase = 0;
    uint32_t cpos = 0;            // ring position of the next unit to consume
    uint32_t accfree_bits = 3;    // per-tile-slot parity of bar_accfree (first wait passes)
    uint32_t pefull_bits = 0;
    auto wrap = [](uint32_t p) -> uint32_t { return p >= (uint32_t)T2_SLOTS ? p - T2_SLOTS : p; };
#define T2_SLOT_WORD(slot) (a_base + (slot) * (T2_SLOT >> 4))
#define T2_WAIT_W()                                                                      \
  do {                                                                                   \
    const uint32_t t_a = mstamp ? (uint32_t)clock64() : 0u;                              \
    if (!P.debug_noload) mbar_wait2(bar_wfull + 8 * wstage, wphase, dbg, 0x400 + wstage); \
    if (mstamp) t_wf += (uint32_t)clock64() - t_a;                                       \
    tc_fence_after();                                                                    \
  } while (0)
#define T2_NEXT_W() do { if (++wstage == T2_WST) { wstage = 0; wphase ^= 1; } } while (0)
#define T2_WE() (share_first ? 0u : bar_wempty + 8 * wstage)
    for (int64_t it = 0; it < n_iter; ++it) {
      for (int op = 0; op < T2_NUM_OPS; ++op) {
        for (int s = 0; s < 2; ++s) {
          const uint32_t d = tmem + (uint32_t)s * TM_ACC;
          mbar_wait2(bar_accfree + 8 * s, (accfree_bits >> s) & 1u, dbg, 0x300 + s);
          accfree_bits ^= 1u << s;
          if (op == 0) {
            mbar_wait2(bar_pefull + 8 * s, (pefull_bits >> s) & 1u, dbg, 0x500 + s);
            pefull_bits ^= 1u << s;
          }
          tc_fence_after();
          const bool mstamp = T2_TIMING && P.timing && blockIdx.x == 0 && it == T2_STAMP_IT && lane == 0;
          const uint32_t t_m0 = mstamp ? (uint32_t)clock64() : 0u;
          uint32_t t_wf = 0;
          const bool share_first = T2_SHARE_SLABS && op >= 7 && s == 0;
          const uint32_t ws_save = wstage, wp_save = wphase;
          if (op <= 6) {
            // forward: units of hi + lo slots, two weight slabs (of 2 k-steps) per unit
            const int n_units = op == 0 ? 1 : (op == 4 ? 5 : 4);
            for (int u = 0; u < n_units; ++u) {
              const uint32_t h = cpos, l = wrap(cpos + 1);
              cpos = wrap(cpos + 2);
              const uint32_t ah = T2_SLOT_WORD(h), al = T2_SLOT_WORD(l);
              T2_WAIT_W();
              issue_unit<128, 2, 3, 256>(d, ah, al, (((128 * 16) >> 4) << 16) | (w_base + wstage * (TC_STAGE_BYTES >> 4)), (uint32_t)(u > 0),
                                         bar_wempty + 8 * wstage, 0u, 0u);
              T2_NEXT_W();
              T2_WAIT_W();
              issue_unit<128, 2, 3, 256>(d, ah + ((2 * T2_KSTEP) >> 4), al + ((2 * T2_KSTEP) >> 4),
                                         (((128 * 16) >> 4) << 16) | (w_base + wstage * (TC_STAGE_BYTES >> 4)), 1u, bar_wempty + 8 * wstage,
                                         bar_afree + 8 * h, bar_afree + 8 * l);
              T2_NEXT_W();
            }
          } else if (op == 7) {
            if (P.rgb3) {   // hi + lo units, one slab per unit
              for (int u = 0; u < 4; ++u) {
                const uint32_t h = cpos, l = wrap(cpos + 1);
                cpos = wrap(cpos + 2);
                T2_WAIT_W();
                issue_unit<64, 4, 3, 128>(d, T2_SLOT_WORD(h), T2_SLOT_WORD(l), (((64 * 16) >> 4) << 16) | (w_base + wstage * (TC_STAGE_BYTES >> 4)),
                                          (uint32_t)(u > 0), T2_WE(), bar_afree + 8 * h, bar_afree + 8 * l);
                T2_NEXT_W();
              }
            } else {        // hi units, two units per slab
              for (int sl = 0; sl < 2; ++sl) {
                const uint32_t h0 = cpos, h1 = wrap(cpos + 1);
                cpos = wrap(cpos + 2);
                T2_WAIT_W();
                const uint32_t bw = (((64 * 16) >> 4) << 16) | (w_base + wstage * (TC_STAGE_BYTES >> 4));
                issue_unit<64, 4, 1, 128>(d, T2_SLOT_WORD(h0), 0u, bw, (uint32_t)(sl > 0), 0u, bar_afree + 8 * h0, 0u);
                issue_unit<64, 4, 1, 128>(d, T2_SLOT_WORD(h1), 0u, bw + ((4 * 2 * 64 * 16) >> 4), 1u, T2_WE(), bar_afree + 8 * h1, 0u);
                T2_NEXT_W();
              }
            }
          } else if (op == 10 || op == 15) {
            // N = 64 products: the whole op is one slab.  Op 10 reads the units that op 11 reads again (and passes): tile slot 0's
            // are the next four in the ring, tile slot 1's the four behind them
            const uint32_t p0 = op == 10 ? wrap(cpos + 4 * s) : cpos;
            if (op == 15) cpos = wrap(cpos + 4);
            T2_WAIT_W();
            const uint32_t bw = (((32 * 16) >> 4) << 16) | (w_base + wstage * (TC_STAGE_BYTES >> 4));
#pragma unroll
            for (int u = 0; u < 4; ++u)
              issue_unit<32, 4, 1, 64>(d, T2_SLOT_WORD(wrap(p0 + u)), 0u, bw + u * ((4 * 2 * 32 * 16) >> 4), (uint32_t)(u > 0),
                                       u == 3 ? T2_WE() : 0u, 0u, 0u);
            T2_NEXT_W();
          } else {
            // backward layer: hi units, one slab per unit; nothing waits for these slots (see the epilogue)
            for (int u = 0; u < 4; ++u) {
              const uint32_t h = cpos;
              cpos = wrap(cpos + 1);
              T2_WAIT_W();
              issue_unit<128, 4, 1, 256>(d, T2_SLOT_WORD(h), 0u, (((128 * 16) >> 4) << 16) | (w_base + wstage * (TC_STAGE_BYTES >> 4)), (uint32_t)(u > 0),
                                         T2_WE(), 0u, 0u);
              T2_NEXT_W();
            }
          }
          if (share_first) { wstage = ws_save; wphase = wp_save; }   // tile slot 1 reads the same slabs (and releases them)
          if (elect_one()) tc_commit2(bar_accfull + 8 * s);
          __syncwarp();
          if (mstamp) P.timing[66 + 2 * op + s] = (long long)(((unsigned long long)(uint32_t)clock64() << 32) | t_m0);
          if (mstamp && s == 0) P.timing[98 + op] = t_wf;
        }
      }
    }
#undef T2_SLOT_WORD
#undef T2_WAIT_W
#undef T2_NEXT_W
#undef T2_WE
  } else {
    // =============================== epilogue warps ===============================
    const int q = warp & 3, sub = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t row_off = (uint32_t)row * 16;
    const uint32_t leader_pefull = mapa_u32(bar_pefull, 0), leader_accfree = mapa_u32(bar_accfree, 0);
    uint32_t ppos = 0;              // ring position of the next unit to produce
    // bar_afree is used only for units that a forward op or the rgb head consumes (ops 0..7): a slot whose previous occupant
    // was read by a backward op (8..15) is rewritten only by phases that have already seen an accumulator whose MMAs were
    // issued after that op, so it needs neither a commit nor a wait
    uint32_t free_bits = 0;         // per-slot parity of bar_afree
    uint32_t need_bits = 0;         // slot's current occupant is consumed by an op <= 7: its next writer waits on bar_afree
    uint32_t accfull_bits = 0;
    uint32_t* const rscr = P.relu_scratch + (size_t)blockIdx.x * (2 * 7 * 2 * 512) + threadIdx.x;
    auto wrap = [](uint32_t p) -> uint32_t { return p >= (uint32_t)T2_SLOTS ? p - T2_SLOTS : p; };
    auto wait_free = [&](uint32_t slot, bool fwd_consumed) {   // fwd_consumed: class of the unit about to be written
      if ((need_bits >> slot) & 1u) {
        mbar_wait_slot(bar_afree + 8 * slot, (free_bits >> slot) & 1u, dbg, 0x600 + slot);
        free_bits ^= 1u << slot;
      }
      need_bits = fwd_consumed ? (need_bits | (1u << slot)) : (need_bits & ~(1u << slot));
    };
    // A unit becomes visible to the tensor core (async proxy) through ONE fence.proxy.async per phase, in front of the
    // accumulator release that tells the MMA warp about the phase's units (a fence per unit is a MEMBAR per unit on the
    // backward chain); the encoding unit of op 0 is announced through bar_pefull and carries its own fence.
    auto publish = [&](uint32_t) {};
    long long* stamp_at = nullptr;   // debug stamps of the current phase (CTA 0, thread 0, first tile pair)
    auto acc_wait = [&](int s) {
      mbar_wait2h(bar_accfull + 8 * s, (accfull_bits >> s) & 1u, dbg, 0x700 + s);
      accfull_bits ^= 1u << s;
      tc_fence_after();
      if (T2_TIMING && stamp_at) stamp_at[0] = clock64();
    };
    auto acc_release = [&](int s) {   // every tcgen05.ld of this phase has completed (tcgen05.wait::ld)
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_accfree + 8 * s);
    };
    // positional encoding of one point as a hi + lo unit (model/dimension_kernel.py:5-35); see mlp_tc.cuh for the column map
    auto produce_pe = [&](float px, float py, float pz, int pe_slot) {   // pe_slot >= 0: operand of op 0 of that tile slot
      const float xs[3] = {px, py, pz};
      float yh[3], yl[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) pe_turns(xs[c], yh[c], yl[c]);
      float v[16];
      float sn, cs;
#define DSN_SC(K, C) pe_sincos(yh[C], yl[C], K, sn, cs)
      if (sub == 0) {
        v[0] = xs[0]; v[1] = xs[1]; v[2] = xs[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) { DSN_SC(0, c); v[3 + c] = sn; v[6 + c] = cs; }
#pragma unroll
        for (int c = 0; c < 3; ++c) { DSN_SC(1, c); v[9 + c] = sn; v[12 + c] = cs; }
        DSN_SC(2, 0); v[15] = sn;
      } else if (sub == 1) {
        DSN_SC(2, 0); v[2] = cs;
        DSN_SC(2, 1); v[0] = sn; v[3] = cs;
        DSN_SC(2, 2); v[1] = sn; v[4] = cs;
#pragma unroll
        for (int c = 0; c < 3; ++c) { DSN_SC(3, c); v[5 + c] = sn; v[8 + c] = cs; }
        DSN_SC(4, 0); v[11] = sn; v[14] = cs;
        DSN_SC(4, 1); v[12] = sn; v[15] = cs;
        DSN_SC(4, 2); v[13] = sn;
      } else if (sub == 2) {
        DSN_SC(4, 2); v[0] = cs;
#pragma unroll
        for (int c = 0; c < 3; ++c) { DSN_SC(5, c); v[1 + c] = sn; v[4 + c] = cs; }
#pragma unroll
        for (int c = 0; c < 3; ++c) { DSN_SC(6, c); v[7 + c] = sn; v[10 + c] = cs; }
#pragma unroll
        for (int c = 0; c < 3; ++c) { DSN_SC(7, c); v[13 + c] = sn; }
      } else {
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
      const uint32_t h = ppos, l = wrap(ppos + 1);
      ppos = wrap(ppos + 2);
      wait_free(h, true);
      wait_free(l, true);
      const uint32_t off = (uint32_t)(2 * sub) * A_CHUNK + row_off;
      *reinterpret_cast<uint4*>(smem + S2_A + h * T2_SLOT + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(smem + S2_A + h * T2_SLOT + off + A_CHUNK) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      *reinterpret_cast<uint4*>(smem + S2_A + l * T2_SLOT + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(smem + S2_A + l * T2_SLOT + off + A_CHUNK) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      if (pe_slot >= 0) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_pefull + 8 * pe_slot);
      }
    };
    // chain rule through the encoding for this thread's own octaves: g = 32 accumulator columns starting at pe_ld0(sub)
    auto chain_rule = [&](const uint32_t (&g)[32], float px, float py, float pz, float (&gx)[3]) {
      const float xs[3] = {px, py, pz};
      float yh[3], yl[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) pe_turns(xs[c], yh[c], yl[c]);
      gx[0] = gx[1] = gx[2] = 0.f;
#define DSN_G(C, C0) __uint_as_float(g[(C) - (C0)])
#define DSN_OCT(K, C0)                                                                   \
  _Pragma("unroll") for (int c = 0; c < 3; ++c) {                                        \
    float sn, cs;                                                                        \
    sincos_turns_fast(yh[c], yl[c], (K), sn, cs);                                        \
    const float gs = DSN_G(3 + 6 * (K) + c, C0), gc = DSN_G(6 + 6 * (K) + c, C0);        \
    gx[c] = fmaf((gs * cs - gc * sn), (float)(1 << (K)), gx[c]);                         \
  }
      if (sub == 0) {
        gx[0] = DSN_G(0, 0); gx[1] = DSN_G(1, 0); gx[2] = DSN_G(2, 0);
        DSN_OCT(0, 0) DSN_OCT(1, 0)
      } else if (sub == 1) {
        DSN_OCT(2, 8) DSN_OCT(3, 8) DSN_OCT(4, 8)
      } else if (sub == 2) {
        DSN_OCT(5, 32) DSN_OCT(6, 32) DSN_OCT(7, 32)
      } else {
        DSN_OCT(8, 32) DSN_OCT(9, 32)
      }
#undef DSN_OCT
#undef DSN_G
    };

    // per-tile state of the two tile slots; `c*` belongs to the slot of the current phase, `o*` to the other one
    float cx = 0.f, cy = 0.f, cz = 0.f, csig = 0.f, cg0 = 0.f, cg1 = 0.f, cg2 = 0.f;
    float ox = 0.f, oy = 0.f, oz = 0.f, osig = 0.f, og0 = 0.f, og1 = 0.f, og2 = 0.f;
    // prologue: encoding units of the first two tiles
    {
      const int64_t i0 = (int64_t)blockIdx.x * TC_TILE + row, i1 = ((int64_t)blockIdx.x + gridDim.x) * TC_TILE + row;
      float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
      if (i0 < n_active) p0 = P.active[i0];
      if (i1 < n_active) p1 = P.active[i1];
      cx = p0.x; cy = p0.y; cz = p0.z;
      ox = p1.x; oy = p1.y; oz = p1.z;
      if (n_iter > 0) { produce_pe(cx, cy, cz, 0); produce_pe(ox, oy, oz, 1); }
    }
    const bool stamp = T2_TIMING && P.timing && blockIdx.x == 0 && threadIdx.x == 0;
    for (int64_t it = 0; it < n_iter; ++it) {
      if (stamp && it == T2_STAMP_IT) P.timing[0] = clock64();
      for (int op = 0; op < T2_NUM_OPS; ++op) {
        for (int s = 0; s < 2; ++s) {
          const int64_t tile = blockIdx.x + (2 * it + s) * (int64_t)gridDim.x;
          const int64_t base = tile * TC_TILE;
          const bool live = base + row < n_active;
          const uint32_t t_acc = t_lane + (uint32_t)s * TM_ACC;
          uint32_t* const rs = rscr + s * (7 * 2 * 512);
          if (T2_TIMING) stamp_at = (stamp && it == T2_STAMP_IT) ? P.timing + 2 + 2 * (2 * op + s) : nullptr;
          if (op <= 6) {
            // ---------- forward layer: bias + ReLU, ReLU bits, fp16 hi / lo units of the next layer's operand.  Layer 6 (density head in
            // fp32, hi-only units for a single-pass rgb head) is a separate instantiation: as predicated code inside the common body it
            // cost every layer 8 loads, 16 descriptor moves and 8 FMAs per unit that do nothing.
            if (op == 3) produce_pe(cx, cy, cz, -1);   // head of layer 4's operand, consumed before this layer's four units
            auto fwd_phase = [&](auto l6_tag) {
              constexpr bool last6 = decltype(l6_tag)::value;
              const float* __restrict__ bias = P.bias + op * 256 + sub * TC_CPT;
              const bool with_lo = !(last6 && !P.rgb3);
              const float* __restrict__ wdp = P.w_dens + sub * TC_CPT;
              float2 sig2 = make_float2(0.f, 0.f);
              acc_wait(s);
              const uint32_t t_accb = t_acc + sub * TC_CPT;
              uint32_t va[16], vb[16];
              tmem_ld16_nowait(t_accb, va);
              uint32_t mw0 = 0, mw1 = 0;
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                uint32_t(&v)[16] = (q4 & 1) ? vb : va;
                float4 b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) b[i] = __ldg(reinterpret_cast<const float4*>(bias + q4 * 64) + i);
                tmem_wait_ld(v);
                if (q4 < 3) { if (q4 & 1) tmem_ld16_nowait(t_accb + (q4 + 1) * 64, va); else tmem_ld16_nowait(t_accb + (q4 + 1) * 64, vb); }
                uint32_t hi[8], lo[8];
                uint32_t mw = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 bb = b[j >> 1];
                  const float2 b2 = (j & 1) ? make_float2(bb.z, bb.w) : make_float2(bb.x, bb.y);
                  const float2 a = __fadd2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), b2);
                  // relu through the fp16 image: hi = relu(fp16(a)) = fp16(relu(a)) (rounding is monotonic), and the mask (hi > 0) that
                  // becomes the ReLU bit also clears the lo part of a non-positive pre-activation: three instructions less per pair
                  // than two fp32 max, bit-identical results
                  const __half2 hr = __floats2half2_rn(a.x, a.y);
                  const uint32_t m2 = __hgt2_mask(hr, as_h2(0u));
                  hi[j] = *reinterpret_cast<const uint32_t*>(&hr) & m2;
                  const int p = j + 8 * (q4 & 1);
                  mw |= m2 & ((1u << p) | (1u << (16 + p)));
                  if (last6) {
                    const float2 wd = __ldg(reinterpret_cast<const float2*>(wdp + q4 * 64 + 2 * j));
                    sig2 = __ffma2_rn(wd, make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f)), sig2);
                  }
                  const float2 hf = __half22float2(hr);
                  const float2 l2 = __ffma2_rn(hf, make_float2(-1.f, -1.f), a);
                  lo[j] = pack_h2(l2.x, l2.y) & m2;
                }
                if (q4 >> 1) mw1 |= mw; else mw0 |= mw;
                const uint32_t hs = ppos;
                const uint32_t off = (uint32_t)(2 * sub) * A_CHUNK + row_off;
                wait_free(hs, true);
                *reinterpret_cast<uint4*>(smem + S2_A + hs * T2_SLOT + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(smem + S2_A + hs * T2_SLOT + off + A_CHUNK) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                if (with_lo) {
                  const uint32_t ls = wrap(ppos + 1);
                  wait_free(ls, true);
                  *reinterpret_cast<uint4*>(smem + S2_A + ls * T2_SLOT + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                  *reinterpret_cast<uint4*>(smem + S2_A + ls * T2_SLOT + off + A_CHUNK) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                  ppos = wrap(ppos + 2);
                } else {
                  ppos = wrap(ppos + 1);
                }
                publish(hs);
              }
              acc_release(s);
              __stcg(rs + (op * 2 + 0) * 512, mw0);
              __stcg(rs + (op * 2 + 1) * 512, mw1);
              if (last6) csig = sig2.x + sig2.y;
            };
            if (op == 6) fwd_phase(std::true_type{}); else fwd_phase(std::false_type{});
          } else if (op == 7) {
            // ---------- rgb head: tail of the rgb layer, sigma / essence out, then the seed units of the backward chain
            // (G6 = (w_dens / scale) * relu'(a6), rebuilt from layer 6's ReLU bits; their loads fly during the tail)
            const uint32_t m0 = __ldcg(rs + (6 * 2 + 0) * 512), m1 = __ldcg(rs + (6 * 2 + 1) * 512);
            const uint4* seedp = reinterpret_cast<const uint4*>(P.seed_h2 + (sub * TC_CPT) / 2);
            uint4 sd0 = __ldg(seedp), sd1 = __ldg(seedp + 1);
            float2 e0 = make_float2(0.f, 0.f), e1 = e0, e2 = e0;
            acc_wait(s);
            const float* rgbw = reinterpret_cast<const float*>(smem + S2_RGBW);
            const uint32_t t_accb = t_acc + sub * TC_CPT;
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
            // sigma and the essence leave the kernel here: the four threads of a row live in warps q, q+4, q+8, q+12 and meet in
            // accumulator columns 128.. of their TMEM lane (the N = 128 product left them unused)
            const float ee0 = e0.x + e0.y, ee1 = e1.x + e1.y, ee2 = e2.x + e2.y;
            if (sub != 0) tmem_st4(t_acc + TM_XCH + 4 * sub, csig, ee0, ee1, ee2);
            tc_fence_before();
            epi_bar();
            if (sub == 0) {
              tc_fence_after();
              uint32_t xv[16];
              tmem_ld16x(t_acc + TM_XCH, xv);   // columns 4..15 hold the partial sums of subs 1..3
              if (live) {
                float sm[4] = {csig, ee0, ee1, ee2};
#pragma unroll
                for (int k = 0; k < 4; ++k) sm[k] += __uint_as_float(xv[4 + k]) + __uint_as_float(xv[8 + k]) + __uint_as_float(xv[12 + k]);
                P.out_a[base + row] = make_float4(sm[0] + P.b_dens, sm[1] + P.b_rgb2[0], sm[2] + P.b_rgb2[1], sm[3] + P.b_rgb2[2]);
              }
            }
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              uint32_t g[8] = {sd0.x, sd0.y, sd0.z, sd0.w, sd1.x, sd1.y, sd1.z, sd1.w};
              if (q4 < 3) { sd0 = __ldg(seedp + (q4 + 1) * 8); sd1 = __ldg(seedp + (q4 + 1) * 8 + 1); }   // next unit's words (64 columns = 8 uint4 further)
              const uint32_t mw = (q4 >> 1) ? m1 : m0;
              if ((q4 & 1) == 0) {
                g[0] &= relu_mask2<0>(mw); g[1] &= relu_mask2<1>(mw); g[2] &= relu_mask2<2>(mw); g[3] &= relu_mask2<3>(mw);
                g[4] &= relu_mask2<4>(mw); g[5] &= relu_mask2<5>(mw); g[6] &= relu_mask2<6>(mw); g[7] &= relu_mask2<7>(mw);
              } else {
                g[0] &= relu_mask2<8>(mw); g[1] &= relu_mask2<9>(mw); g[2] &= relu_mask2<10>(mw); g[3] &= relu_mask2<11>(mw);
                g[4] &= relu_mask2<12>(mw); g[5] &= relu_mask2<13>(mw); g[6] &= relu_mask2<14>(mw); g[7] &= relu_mask2<15>(mw);
              }
              const uint32_t hs = ppos;
              ppos = wrap(ppos + 1);
              const uint32_t off = (uint32_t)(2 * sub) * A_CHUNK + row_off;
              wait_free(hs, false);
              *reinterpret_cast<uint4*>(smem + S2_A + hs * T2_SLOT + off) = make_uint4(g[0], g[1], g[2], g[3]);
              *reinterpret_cast<uint4*>(smem + S2_A + hs * T2_SLOT + off + A_CHUNK) = make_uint4(g[4], g[5], g[6], g[7]);
              publish(hs);
            }
            acc_release(s);
          } else if (op == 10) {
            // ---------- d sigma / d PE through layer 4: chain rule at once, three partial sums survive
            acc_wait(s);
            uint32_t g[32];
            tmem_ld32(t_acc + pe_ld0(sub), g);
            acc_release(s);
            float gx[3];
            chain_rule(g, cx, cy, cz, gx);
            cg0 = gx[0] * P.stash_scale; cg1 = gx[1] * P.stash_scale; cg2 = gx[2] * P.stash_scale;
          } else if (op == 15) {
            // ---------- last op: the next tile's encoding unit first (the MMA warp can start on it), then the gradient output
            const bool more = it + 1 < n_iter;
            float4 pn = make_float4(0.f, 0.f, 0.f, 0.f);
            if (more) {
              const int64_t nb = (tile + 2 * (int64_t)gridDim.x) * TC_TILE + row;
              if (nb < n_active) pn = P.active[nb];
              produce_pe(pn.x, pn.y, pn.z, s);
            }
            acc_wait(s);
            uint32_t g[32];
            tmem_ld32(t_acc + pe_ld0(sub), g);
            float gx[3];
            chain_rule(g, cx, cy, cz, gx);
            gx[0] += cg0; gx[1] += cg1; gx[2] += cg2;
            if (sub != 0) tmem_st4(t_acc + TM_XCH + 4 * sub, gx[0], gx[1], gx[2], 0.f);
            tc_fence_before();
            epi_bar();
            if (sub == 0) {
              tc_fence_after();
              uint32_t xv[16];
              tmem_ld16x(t_acc + TM_XCH, xv);
              if (live) {
#pragma unroll
                for (int k = 0; k < 3; ++k) gx[k] += __uint_as_float(xv[4 + k]) + __uint_as_float(xv[8 + k]) + __uint_as_float(xv[12 + k]);
                P.out_g[base + row] = make_float4(gx[0] * P.seed_scale, gx[1] * P.seed_scale, gx[2] * P.seed_scale, 0.f);
              }
            }
            acc_release(s);
            cx = pn.x; cy = pn.y; cz = pn.z;
            if (stamp && it == T2_STAMP_IT && s == 1) P.timing[1] = clock64();
          } else {
            // ---------- backward layer: G_{l-1} = (G_l W_l) * relu'(a_{l-1}), fp16 hi-only units
            const int layer = op == 8 ? 5 : (op == 9 ? 4 : 14 - op);   // ops 11..14 -> layers 3..0
            const uint32_t m0 = __ldcg(rs + (layer * 2 + 0) * 512), m1 = __ldcg(rs + (layer * 2 + 1) * 512);
            acc_wait(s);
            const uint32_t t_accb = t_acc + sub * TC_CPT;
            uint32_t va[16], vb[16];
            tmem_ld16_nowait(t_accb, va);
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
              const uint32_t hs = ppos;
              ppos = wrap(ppos + 1);
              const uint32_t off = (uint32_t)(2 * sub) * A_CHUNK + row_off;
              wait_free(hs, false);
              *reinterpret_cast<uint4*>(smem + S2_A + hs * T2_SLOT + off) = make_uint4(g[0], g[1], g[2], g[3]);
              *reinterpret_cast<uint4*>(smem + S2_A + hs * T2_SLOT + off + A_CHUNK) = make_uint4(g[4], g[5], g[6], g[7]);
              publish(hs);
            }
            acc_release(s);
          }
          if (T2_TIMING && stamp_at) stamp_at[1] = clock64();
          // the next phase belongs to the other tile slot
          { float t;
            t = cx; cx = ox; ox = t; t = cy; cy = oy; oy = t; t = cz; cz = oz; oz = t; t = csig; csig = osig; osig = t;
            t = cg0; cg0 = og0; og0 = t; t = cg1; cg1 = og1; og1 = t; t = cg2; cg2 = og2; og2 = t; }
        }
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
  cluster_sync_all();
  if (warp == TC_EPI_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
  }
}

inline void tc2_configure() { cudaFuncSetAttribute(mlp_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2_SMEM); }

inline size_t tc2_scratch_bytes(int sm_count) { return (size_t)(sm_count & ~1) * 2 * 7 * 2 * 512 * sizeof(uint32_t); }

inline int tc2_launch(TcWeights& w, long long* timing, unsigned int* dbg, int debug_noload, uint32_t* relu_scratch, const float4* active,
                      const unsigned long long* n_active_ptr, int64_t n_active_host, float4* out_a, float4* out_g, int sm_count, cudaStream_t st) {
  Tc2Params p{};
  for (int i = 0; i < T2_NUM_OPS; ++i) p.ops[i] = w.ops2[i];
  p.wpack = reinterpret_cast<const uint8_t*>(w.d_pack2);
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
  p.relu_scratch = relu_scratch;
  p.rgb3 = w.rgb3 ? 1 : 0;
  p.timing = timing;
  p.dbg = dbg;
  p.debug_noload = debug_noload;
  mlp_tc2_kernel<<<sm_count & ~1, TC_THREADS, T2_SMEM, st>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace dsn

// Flash-style prefill attention on tcgen05 tensor cores (sm_100a): S = Q K^T and O += P V are tcgen05.mma with
// the accumulators in TMEM; K/V/Q tiles arrive by TMA (64B-swizzled atoms of 32 head-dim columns, so head_dim 96
// needs no padding); softmax runs thread-per-row out of TMEM.
//
// Product configurations (end of round 2; DESIGN.md 6a-6c has the measurements behind every choice):
//   head_dim 64 / 96 (NT = 1): one CTA = one 128-row query tile of one head at a time, 256 threads, 256 TMEM columns
//     and 48 KB (hd 64, P through TMEM) / 112 KB (hd 96, P through 128B-swizzled smem, row sums from 16 ones columns
//     of V) of shared memory, so TWO CTAs share an SM; a CTA walks several query tiles (causal: the pair {nt-1-x, x}).
//   head_dim 128 (NT = 2): one CTA per SM, two query tiles (A, B) sharing a 1-stage K/V ring, all 512 TMEM columns,
//     row sums in registers, the MMA warps converged with an elected issuing lane.
// Roles:  warp 0 TMA producer | warp 1 (and 3 for tile B) MMA issuer | warp 2 TMEM allocator | warps 4.. softmax:
//   tcgen05.ld the whole 128-column S row into registers and hand the S buffer straight back, so S(j+1) is computed
//   while this block's softmax runs -> chunk-level masking (only blocks that need it) -> running max -> packed FFMA2
//   scale-and-subtract + exp2 -> bf16 P; the first 64 keys are published on their own barrier so that half of the
//   P.V MMAs are issued under the remaining exponentials; O is rescaled in TMEM lazily, only when the row max grew by
//   more than 2^8.
// The LR_ATTN_* switches below are the experiments of rounds 1 and 2. Product defaults: P_TMEM, P_HALF, EXP_FIRST,
// DESC32, CHUNK_MASK, FFMA2 = 1, MMA_WARP = 2; everything else is 0 = measured neutral or slower (each comment says
// by how much; tools/attn_ab.py builds any combination and times it interleaved against the product).
#include <cuda.h>

#include <atomic>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap_cache.cuh"

namespace lr {

// Optional in-kernel timeline (tools/attn_trace.py builds a separate .so with -DLR_ATTN_TRACE): clock64 stamps of
// CTA (0,0,0) at the main pipeline events, [role][event][block] -> global buffer.
#ifdef LR_ATTN_TRACE
// per-CTA record {SM id, start ns, end ns, K/V blocks} of EVERY CTA (tools/attn_cta_times.py): occupancy timeline
__device__ long long* g_attn_cta = nullptr;
__device__ __forceinline__ long long attn_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ long long* g_attn_trace = nullptr;
#define ATTN_TRACE(role, ev, j)                                                                     \
  do {                                                                                              \
    if (g_attn_trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 32)          \
      g_attn_trace[((role) * 8 + (ev)) * 32 + (j)] = clock64();                                     \
  } while (0)
#else
#define ATTN_TRACE(role, ev, j) \
  do {                          \
  } while (0)
#endif

constexpr int kAtomBytes = 128 * 64;      // [128 rows x 32 bf16], 64B swizzle
constexpr int kPBytes = 128 * 128 * 2;    // P tile: two [128 x 64 bf16] 128B-swizzled atoms
#ifndef LR_ATTN_NT2_STAGGER
#define LR_ATTN_NT2_STAGGER 1100
#endif
constexpr long long kStaggerCycles = LR_ATTN_NT2_STAGGER;
constexpr float kRescaleThreshold = 8.f;  // log2 units: P stays <= 2^8 between rescales

__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= uint64_t((addr >> 4) & 0x3FFF);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(layout) << 61;
  return d;
}
constexpr uint32_t kLayoutSW128 = 2, kLayoutSW64 = 4;
// the same descriptor as two 32-bit words: lo = start address | leading byte offset, hi = stride byte offset | version
// | layout (a compile-time constant per operand kind); a byte offset inside the tile is one add of (offset >> 4) to lo
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t addr, uint32_t lbo_bytes) {
  return ((addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ constexpr uint32_t umma_desc_hi(uint32_t sbo_bytes, uint32_t layout) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
}
__device__ __forceinline__ uint64_t desc_join(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// Softmax exponentials: LR_ATTN_POLY_NUM of every 4 elements of a row are computed on the FMA pipe instead of the
// MUFU (16 ex2/clk/SM is the scarcest pipe of this kernel): Cody-Waite split 2^x = 2^floor(x) * 2^frac with a
// round-down add against 1.5 * 2^23 (floor(x) lands in the low mantissa bits), a degree-3 minimax polynomial for
// 2^frac on [0, 1) (max relative error 9e-5, well under the bf16 rounding of P) and an integer add into the
// exponent field. Inputs are clamped to -127, so masked (-inf) entries become ~2^-127 instead of 0; rows that are
// masked completely are zero-filled by the epilogue either way.
#ifndef LR_ATTN_POLY_NUM
#define LR_ATTN_POLY_NUM 0
#endif
// Release the S buffer as soon as the row is in registers (before mask + max) when one thread owns the whole row.
#ifndef LR_ATTN_EARLY_SFREE
#define LR_ATTN_EARLY_SFREE 0
#endif
// Experiments on the MMA-issue path (default off = the measured product; tools/attn_variants.py builds them):
// LR_ATTN_HOIST_DESC 1 builds the shared-memory descriptors once and adds the byte offset (>> 4) per tcgen05.mma instead
// of rebuilding both descriptors from the address every time; LR_ATTN_AUX_REGS is the register count the non-softmax
// warps keep after setmaxnreg in the one-tile configuration (40, or 48 = 128 x 48 + 128 x 208 = the CTA's allocation).
#ifndef LR_ATTN_HOIST_DESC
#define LR_ATTN_HOIST_DESC 0
#endif
#ifndef LR_ATTN_AUX_REGS
#define LR_ATTN_AUX_REGS 40
#endif
// Softmax-side experiments of round 2 (default = the measured best; tools/attn_variants.py builds and times them):
// LR_ATTN_PIPE_LD 1 loads the S row from TMEM in four 32-column chunks and folds the (mask +) running max of chunk c
// under the tcgen05.ld of chunk c+1 instead of waiting for the whole row first; LR_ATTN_MAX3 1 uses the three-input
// FMNMX3 (max.f32 d, a, b, c) - 64 instead of 128 ALU-pipe instructions per row.
#ifndef LR_ATTN_PIPE_LD
#define LR_ATTN_PIPE_LD 0
#endif
// LR_ATTN_P_TMEM 1 (head_dim 64, one tile per CTA): P never touches shared memory - the softmax threads write the bf16
// P row into 64 TMEM columns of their own (tcgen05.st) and the P.V MMA takes its A operand from TMEM
// (tcgen05.mma ... [d], [a_tmem], b_desc); S 128 + P 64 + O 64 = the CTA's 256 columns, so the row sums move to
// registers (no ones columns), the P smem buffer, its 16 STS.128 per row and the async-proxy fence disappear and the
// CTA's shared memory drops from 88 KB to 48 KB. head_dim 96 does not fit (128 + 64 + 96 > 256).
#ifndef LR_ATTN_P_TMEM
#define LR_ATTN_P_TMEM 1   // product since r02: CLIP attention 1.265 -> 1.199 ms (profiles/r02_attention_knockouts.txt)
#endif
// Knock-out experiments (WRONG results, timing only; tools/attn_variants.py --define=LR_ATTN_KO=n): bit 0 replaces the
// MUFU exponential by one FMUL, bit 1 skips the P stores + the async-proxy fence, bit 2 skips the running max,
// bit 4 skips the P.V MMAs (commits only), bit 5 skips the S MMAs, bit 6 keeps the P stores but drops the async-proxy
// fence, bit 7 loads K / V by TMA only for the first block of a tile (later blocks reuse the stale tile: no L2 traffic). They show which resource the product kernel's time
// is actually sensitive to (profiles/r02_attention_knockouts.txt).
#ifndef LR_ATTN_KO
#define LR_ATTN_KO 0
#endif
#ifndef LR_ATTN_MAX3
#define LR_ATTN_MAX3 0
#endif
// Hand-off experiments of the second half of round 2 (tools/attn_ab.py, interleaved A/B timing):
// LR_ATTN_P_HALF 1: the softmax warps arrive on a second barrier (p_half) once the first 64 keys of the P row are
// stored, and the MMA thread issues the first four P.V MMAs then - half of the P.V issue latency moves under the
// exponentials of the second half instead of sitting between p_full and pv_done.
// LR_ATTN_EXP_FIRST 1: the exponentials of the first 32 keys are computed BEFORE the wait on pv_done(j-1) (they only
// need registers), so that wait overlaps useful work.
// LR_ATTN_DESC32 1: the MMA thread builds every shared-memory descriptor as {base_lo + constant, constant hi} with one
// 32-bit add instead of masking / shifting the address per tcgen05.mma (4 live registers instead of 28 descriptors).
// LR_ATTN_KV2 1 (head_dim 64 with P in TMEM: 80 KB per CTA): two-stage K / V ring with two CTAs per SM.
#ifndef LR_ATTN_P_HALF
#define LR_ATTN_P_HALF 1
#endif
#ifndef LR_ATTN_EXP_FIRST
#define LR_ATTN_EXP_FIRST 1
#endif
#ifndef LR_ATTN_DESC32
#define LR_ATTN_DESC32 1
#endif
#ifndef LR_ATTN_KV2
#define LR_ATTN_KV2 0
#endif
// LR_ATTN_STAGGER = D > 0 (two CTAs per SM): the SECOND CTA that lands on an SM starts its pipeline D cycles late.
// Two co-resident CTAs that start together run their MUFU-bound exponential phases at the same time, each at half
// rate, and - the phase offset of two equal loops that share one pipe is neutrally stable - stay that way for the
// whole kernel; an initial offset of about one exponential phase lets each CTA's exponentials run under the other's
// S load / max / hand-offs. Mode 1: "second" = linear CTA index in [#SMs, 2 #SMs) (the hardware fills the SMs
// breadth-first); mode 2: an atomic per-SM arrival counter tagged with a per-launch epoch.
// LR_ATTN_MMA_WARP 1: the MMA-issuing warp stays CONVERGED - all 32 lanes walk the event loop (barrier probes are
// made warp-uniform with a vote) and one elected lane (elect.sync) issues tcgen05.mma / tcgen05.commit. Under
// `if (lane == 0)` the compiler cannot know that a single lane is active and wraps every UTCHMMA in an ELECT /
// BRA.U.ANY loop, with the descriptors moved from vector to uniform registers (R2UR) each time.
// 0 = off, 1 = every configuration, 2 = only with two query tiles per CTA (one CTA per SM), where it measured faster.
#ifndef LR_ATTN_MMA_WARP
#define LR_ATTN_MMA_WARP 2
#endif
// kMmaWarp is a constexpr bool of the kernel (the macro value resolved against NT)
#define LR_MMA_VOTE(x) (kMmaWarp ? __all_sync(0xffffffffu, (x)) : (x))
#define LR_MMA_LEAD_BEGIN if (!kMmaWarp || elect_one()) {
#define LR_MMA_LEAD_END            \
  }                                \
  if constexpr (kMmaWarp) __syncwarp();
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
// LR_ATTN_CHUNK_MASK 1: in a block that needs masking (causal diagonal, last partial block) the valid keys of a row are
// a prefix [0, nvalid) of the block's 128 columns. The warp classifies each 32-column chunk from the min / max of nvalid
// over its 32 rows (redux.sync): valid for every lane -> no masking instructions; masked for every lane -> no max, no
// exponentials (P = 0 is stored directly); only the one or two chunks in between pay a per-element compare + select
// (one ISETP + SEL against nvalid instead of two compares). Bit-identical results: exp2(-inf) is exactly 0 either way.
#ifndef LR_ATTN_CHUNK_MASK
#define LR_ATTN_CHUNK_MASK 1   // product: decoder -7 %, CLIP -3 % (profiles/r02_attention_ab_interleaved_6_chunk_mask.txt)
#endif
// LR_ATTN_SPIN_WAIT 1: the softmax warps poll s_full / pv_done with mbarrier.test_wait instead of try_wait (which may
// suspend the warp until the phase completes or a time slice ends).
// LR_ATTN_PAD_SMEM = bytes of unused shared memory added to the two-CTAs-per-SM configurations: with 100 KB only ONE
// CTA fits on an SM - the measurement of what co-residency is worth.
#ifndef LR_ATTN_PAD_SMEM
#define LR_ATTN_PAD_SMEM 0
#endif
// LR_ATTN_SPEC_MAX 1 (one softmax thread per row, blocks after the first of a tile that need no masking): the
// exponentials start with the OLD reference maximum while the row maximum of this block is reduced in the issue slots
// the MUFU-paced exponentials leave idle; the first 64 keys of P are held in registers until the maximum is known.
// If no row of the warp has outgrown the lazy-rescale threshold (the usual case after the first blocks) the values are
// exactly those of the plain path; otherwise the warp rescales O and recomputes those 64 keys with the new reference -
// also exactly the plain path's values. The S buffer goes back to the MMA thread right after the TMEM read.
// Measured: bit-identical output, +6 % slower on every shape (profiles/r02_attention_ab_interleaved_16_speculative_max.txt). Off.
#ifndef LR_ATTN_SPEC_MAX
#define LR_ATTN_SPEC_MAX 0
#endif
// LR_ATTN_P_HYBRID 1 (head_dim 96, one tile per CTA): the decoder shape is shared-memory-bandwidth bound and
// TMEM has 256 - 128 (S) - 96 (O) = 32 spare columns once the row sums live in registers: exactly the first 64 keys of
// the bf16 P row. That half goes through TMEM (tcgen05.st, A operand of the first four P.V MMAs from TMEM), the second
// half through shared memory as before: 36 KB less shared-memory traffic per block (16 KB of P stores, 16 KB of P
// operand reads, the 4 KB ones atom of V).
#ifndef LR_ATTN_P_HYBRID
#define LR_ATTN_P_HYBRID 0
#endif
// LR_ATTN_NO_ONES 1: no [V | 1] row-sum columns anywhere - the softmax threads sum the exponentials in registers (as the
// P-in-TMEM and two-tile hd-128 configurations already do): 4 KB less operand traffic per block, 128 FADD more per row.
#ifndef LR_ATTN_NO_ONES
#define LR_ATTN_NO_ONES 0
#endif
#ifndef LR_ATTN_SPIN_WAIT
#define LR_ATTN_SPIN_WAIT 0
#endif
// LR_ATTN_PRELOAD 1 (one softmax thread per row): software pipelining inside the softmax thread. While the
// exponentials of block j run, the registers of the chunks already consumed are refilled with S(j+1) - probed with a
// non-blocking test of s_full once the first / second chunk is done, so nothing stalls if the S MMA is late - and after
// the p_full arrive of block j only `tcgen05.wait::ld` + the row max of block j+1 are left before s_free(j+1). The wait
// on s_full and the TMEM read latency (two of the seven links of the per-block chain) leave the critical
// path; blocks that need masking (the diagonal / last block) and the first block of a tile take the plain path.
// Measured (profiles/r02_attention_ab_interleaved_12_preload.txt): +10 ... +26 % SLOWER - the probes, the volatile
// tcgen05.ld between the chunks and the branches cut the exponential phase into basic blocks the scheduler can no
// longer interleave, and S(j+1) is rarely complete when the first two chunks are done. Off.
#ifndef LR_ATTN_PRELOAD
#define LR_ATTN_PRELOAD 0
#endif
// LR_ATTN_L2_PREFETCH 1: with the single-stage K / V ring the load of block j+1 starts only when the MMAs on block j
// have retired, so its latency sits between two S MMAs of a tile; the producer therefore asks the L2 for block j+1
// (cp.async.bulk.prefetch.tensor) while it issues the load of block j.
#ifndef LR_ATTN_L2_PREFETCH
#define LR_ATTN_L2_PREFETCH 0
#endif
__device__ __forceinline__ void tma_prefetch_l2_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(c0), "r"(c1)
               : "memory");
}
// LR_ATTN_POLL_SLEEP = N > 0: the two event loops (TMA producer, MMA issuer) sleep N ns (nanosleep) after an iteration
// in which nothing was ready. Their mbarrier.test_wait polling otherwise issues continuously on schedulers 0 and 1,
// which the softmax warps of BOTH co-resident CTAs share.
#ifndef LR_ATTN_POLL_SLEEP
#define LR_ATTN_POLL_SLEEP 0
#endif
__device__ __forceinline__ void poll_idle() {
#if LR_ATTN_POLL_SLEEP > 0
  asm volatile("nanosleep.u32 %0;" ::"n"(LR_ATTN_POLL_SLEEP));
#endif
}
// LR_ATTN_EPI_STAGE 1: the O tile leaves through shared memory. A softmax thread owns one output ROW, so its direct
// 16-byte stores hit 32 different 128-byte lines per warp instruction - the timeline shows 1300 (head_dim 64) to 3800
// (head_dim 96) cycles per tile in the store loop, LSU-transaction-bound. Staged: every warp writes its 32 rows into
// a private smem region (its own rows of the retired P buffer, or a dedicated buffer when P lives in TMEM), then
// reads them back so that consecutive lanes carry consecutive 16-byte pieces of a row (3-6 lines per instruction).
#ifndef LR_ATTN_EPI_STAGE
#define LR_ATTN_EPI_STAGE 0
#endif
// LR_ATTN_FFMA2 1: the scale-and-subtract in front of the exponentials as packed fma.rn.f32x2 (two elements per
// instruction: 64 instead of 128 FMA-pipe issues per row and block).
#ifndef LR_ATTN_FFMA2
#define LR_ATTN_FFMA2 1   // product: hd64 -2 %, hd96 -2.6 %, hd128 +-0 (profiles/r02_attention_ab_interleaved_7_ffma2.txt)
#endif
__device__ __forceinline__ void ffma2(float& d0, float& d1, uint32_t a0, uint32_t a1, float b, float c) {
  uint64_t a, bb, cc, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a0), "r"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(bb), "l"(cc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ void softmax_wait(uint64_t* bar, uint32_t parity) {
#if LR_ATTN_SPIN_WAIT
  while (!mbar_test_wait(bar, parity)) {
  }
#else
  mbar_wait(bar, parity);
#endif
}
#ifndef LR_ATTN_STAGGER
#define LR_ATTN_STAGGER 0
#endif
#ifndef LR_ATTN_STAGGER_MODE
#define LR_ATTN_STAGGER_MODE 1
#endif
#if LR_ATTN_STAGGER > 0 && LR_ATTN_STAGGER_MODE == 2
__device__ unsigned long long g_attn_sm_arrivals[1024];   // [%smid] = epoch << 32 | CTAs of that launch seen so far
#endif
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// tcgen05.wait::ld that names the registers of the load it completes as in/out operands: the compiler may not move
// their consumers above the wait nor the wait above the load (needed once loads and uses are interleaved).
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ float exp2_fma_pipe(float x) {
  x = fmaxf(x, -127.f);
  float r;
  asm("add.rm.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(12582912.f));
  const float f = x - (r - 12582912.f);
  float p = fmaf(f, 0.077119089663028717f, 0.227564394474029541f);
  p = fmaf(p, f, 0.695146143436431885f);
  p = fmaf(p, f, 1.f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}
__device__ __forceinline__ float softmax_exp2(float x, int idx) {  // idx is a constant after unrolling
#if LR_ATTN_KO & 1
  return x * 0.001f;
#else
  return ((idx & 3) < LR_ATTN_POLY_NUM) ? exp2_fma_pipe(x) : exp2f(x);
#endif
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x (K/2) 32-bit columns of packed bf16 pairs (K-major), one CTA.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// NT = query tiles per CTA. NT = 2: one CTA per SM, K/V shared by both tiles, 2-stage rings.
// NT = 1: one tile per CTA, single-stage K/V, 112 KB of smem and 256 TMEM columns -> TWO CTAs per SM, which de-phases
// the two softmax warpgroups of an SM naturally and hides each CTA's prologue/epilogue behind the other's main loop.
template <int HD, int NT>
struct AttnTcCfg {
  // head_dim 128 (Llama decoder of the LLaVA-v1.6 branch): S (128) + O (128 + 16) columns exceed 256 and one CTA
  // needs 136 KB of smem, so it runs one CTA per SM with the 2-stage K/V ring and all 512 TMEM columns.
  static constexpr bool kOnePerSm = NT == 2 || HD > 96;
  // head_dim 128 with two tiles: S_A, S_B, O_A, O_B take all 512 TMEM columns, so there is no room for the 16 row-sum
  // columns of the ones trick - the softmax threads keep the row sums in registers instead (fp32 sum of the un-rounded
  // exponentials, like flash-attention 2) - and 227 KB of smem allow only a single K/V stage.
  static constexpr bool kPTmem = LR_ATTN_P_TMEM && HD == 64 && NT == 1;
  static constexpr bool kPHybrid = LR_ATTN_P_HYBRID && HD == 96 && NT == 1;
  static constexpr bool kOnes = !(HD == 128 && NT == 2) && !kPTmem && !kPHybrid && !LR_ATTN_NO_ONES;
  static constexpr int kStages = !kOnePerSm ? ((kPTmem && LR_ATTN_KV2) ? 2 : 1) : (kOnes ? 2 : 1);
  static constexpr int kTmemCols = kOnePerSm ? 512 : 256;
  static constexpr int kAtoms = HD / 32;
  static constexpr int kTileBytes = kAtoms * kAtomBytes;         // one Q / K tile, and the TMA-loaded part of a V tile
  static constexpr int kVTileBytes = (kAtoms + (kOnes ? 1 : 0)) * kAtomBytes;  // V tile + one atom of ones (row sums via the MMA)
  // epilogue staging (LR_ATTN_EPI_STAGE): row pitch HD*2 + 16 bytes (conflict-free 16-byte accesses), 256 with an XOR
  // swizzle for head_dim 128; with P in TMEM there is no retired P buffer to borrow, so 128 rows are added
  static constexpr int kStagePitch = HD == 128 ? 256 : HD * 2 + 16;
  static constexpr int kStageBytes = (LR_ATTN_EPI_STAGE && kPTmem) ? 128 * kStagePitch : 0;
  static constexpr int kSmemBytes = NT * kTileBytes /*Q*/ + kStages * kTileBytes /*K*/ + kStages * kVTileBytes /*V*/ +
                                    (kPTmem ? 0 : NT * kPBytes) + 256 /*barriers*/ + (NT == 2 ? 2048 : 0) /*row-max exchange*/ +
                                    kStageBytes + (kOnePerSm ? 0 : LR_ATTN_PAD_SMEM);
};

// SPLIT = threads per query row in the softmax: 1 -> one thread owns the whole 128-column S row (2 warpgroups,
// the product path), 2 -> two threads own 64 columns each (4 warpgroups, row maxima exchanged through shared memory).
// The split variant was built to test whether more warps per scheduler help; it is ~15 % SLOWER: per pair of tiles
// and K/V block the kernel moves ~344 KB through shared memory (Q, K, P, V operand reads of the MMAs + P stores +
// TMA fills) = ~2700 cycles at 128 B/clk, which - not MUFU, not latency - bounds the iteration (r01 finding; the next
// step is P through TMEM, aliasing S, as the A operand of the PV MMA).
template <int HD, bool CAUSAL, int SPLIT, int NT>
__global__ void __launch_bounds__(128 + 128 * NT * SPLIT, AttnTcCfg<HD, NT>::kOnePerSm ? 1 : 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, bf16* __restrict__ o, int ld_o, int rows_per_seq,
               const int* __restrict__ seq_base, const int* __restrict__ seq_start, const int* __restrict__ seq_len,
               int q_col0, int k_col0, int v_col0, int kv_group, float scale_log2,
               const int* __restrict__ row_lo, const int* __restrict__ row_hi, int tile_mode, unsigned epoch) {
  using Cfg = AttnTcCfg<HD, NT>;
  constexpr int NS = Cfg::kStages;
  constexpr int NA = Cfg::kAtoms;
  constexpr int TILE = Cfg::kTileBytes;
  constexpr int VTILE = Cfg::kVTileBytes;
  constexpr bool ONES = Cfg::kOnes;
  constexpr int HDX = HD + (ONES ? 16 : 0);  // O columns incl. the row-sum columns
  extern __shared__ __align__(1024) uint8_t smem_raw[];  // the swizzled tiles need a 1024 B aligned base
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  uint8_t* smem = smem_raw;
  uint8_t* sQ = smem;                    // [NT][TILE]
  uint8_t* sK = sQ + NT * TILE;          // [NS][TILE]
  uint8_t* sV = sK + NS * TILE;          // [NS][VTILE]
  uint8_t* sP = sV + NS * VTILE;         // [NT][kPBytes]
  constexpr bool PTMEM = Cfg::kPTmem;
  constexpr bool PHYB = Cfg::kPHybrid;   // keys 0..63 of P in TMEM, 64..127 in the second smem atom
  constexpr bool kMmaWarp = LR_ATTN_MMA_WARP == 1 || (LR_ATTN_MMA_WARP == 2 && NT == 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (PTMEM ? 0 : NT * kPBytes));
  uint64_t* q_full = bars;               // [1]
  uint64_t* k_full = bars + 1;           // [2]
  uint64_t* k_empty = bars + 3;          // [2]
  uint64_t* v_full = bars + 5;           // [2]
  uint64_t* v_empty = bars + 7;          // [2]
  uint64_t* s_full = bars + 9;           // [2] per tile  MMA -> softmax : S_x(j) is in TMEM
  uint64_t* s_free = bars + 11;          // [2] per tile  softmax -> MMA : S_x(j) has been read into registers
  uint64_t* p_full = bars + 13;          // [2] per tile  softmax -> MMA : P_x(j) is in smem (and O_x rescaled)
  uint64_t* pv_done = bars + 15;         // [2] per tile  MMA -> softmax : O_x += P_x(j) V_j retired
  uint64_t* o_final = bars + 17;         // [2] per tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
  uint64_t* p_half = bars + 20;          // [2] per tile  softmax -> MMA : the first 64 keys of P_x(j) are stored (LR_ATTN_P_HALF)
  float* mbuf = reinterpret_cast<float*>(bars + 32);  // [2 tiles][2 halves][128] row-max exchange (SPLIT == 2)
  [[maybe_unused]] uint8_t* sStage = reinterpret_cast<uint8_t*>(bars) + 256 + (NT == 2 ? 2048 : 0);  // PTMEM only
  constexpr int kThreads = 128 + 128 * NT * SPLIT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int kv_head = head / kv_group;  // grouped-query attention: kv_group query heads share one K/V head
  // Query tiles of this CTA (the per-CTA fixed cost - launch, TMEM allocation, barrier setup, pipeline fill and drain,
  // 6.4 us on the causal decoder shape against 2.2 us per K/V block - is paid once per CTA, not once per tile):
  //   tile_mode 0: the one tile blockIdx.x names (causal: heaviest first);
  //   tile_mode 1 (causal): the pair {nt-1-x, x} - every CTA walks nt+1 K/V blocks, so the grid is balanced as well;
  //   tile_mode k >= 2: the k consecutive tiles x*k .. x*k+k-1.
  // Barrier phases run on counters of K/V blocks / tiles processed by this CTA, so they carry across tiles.
  const int nt_total = (rows_per_seq + 128 * NT - 1) / (128 * NT);
  int n_tiles = 1;
  if (tile_mode == 1) n_tiles = (int(blockIdx.x) < nt_total - 1 - int(blockIdx.x)) ? 2 : 1;
  else if (tile_mode >= 2) n_tiles = min(tile_mode, nt_total - int(blockIdx.x) * tile_mode);
  auto tile_qt = [&](int ti) -> int {
    if (tile_mode == 0) return CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
    if (tile_mode == 1) return ti == 0 ? nt_total - 1 - int(blockIdx.x) : int(blockIdx.x);
    return int(blockIdx.x) * tile_mode + ti;
  };
  int m0 = tile_qt(0) * 128 * NT;
  // slot layout: sequence s owns rows [s*rows_per_seq, +rows_per_seq), valid run [start, start+len);
  // packed layout (seq_base != NULL): sequence s owns exactly rows [seq_base[s], +seq_len[s]), rows_per_seq = max len
  // segment layout (row_lo != NULL; non-causal, NT == 1): ONE buffer of rows_per_seq rows cut into short segments (the
  // <= 64-token windows of the Qwen2.5-VL vision tower); row r attends to keys [row_lo[r], row_hi[r]). A CTA takes 128
  // consecutive rows - several segments - and the K/V rows from the first row's segment start to the last row's
  // segment end (<= 2 blocks of 128), masking per row: full 128-row tiles instead of one half-empty tile per window.
  const bool seg = row_lo != nullptr;
  int start = (seq_start && !seq_base) ? seq_start[seq] : 0;
  const int len = seq_len ? seq_len[seq] : rows_per_seq;
  int end = start + len;
  const int slot_row0 = seq_base ? seq_base[seq] : seq * rows_per_seq;
  if (seg) {
    start = row_lo[m0];
    end = row_hi[min(m0 + 128 * NT, rows_per_seq) - 1];
  }

  int lo[2], hi[2], nblk[2], kv_end[2], n;
  auto set_tile = [&](int ti) {   // every role calls this with the same ti sequence
    m0 = tile_qt(ti) * 128 * NT;
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      lo[x] = max(m0 + x * 128, start);
      hi[x] = min(min(m0 + x * 128 + 128, end), rows_per_seq);
      const bool valid = lo[x] < hi[x] && x < NT;
      kv_end[x] = valid ? (CAUSAL ? min(end, hi[x]) : end) : start;
      nblk[x] = (kv_end[x] - start + 127) / 128;
    }
    n = max(nblk[0], nblk[1]);
  };
  set_tile(0);
#ifdef LR_ATTN_TRACE
  long long cta_t0 = 0;
  if (threadIdx.x == 0) cta_t0 = attn_globaltimer();
#endif

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_qkv);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], NT);  // one arrival per MMA-issuing thread (one per tile)
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], NT);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4 * SPLIT);
      mbar_init(&p_full[i], 4 * SPLIT);
      mbar_init(&p_half[i], 4 * SPLIT);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_final[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  // the ones atom of both V stages (bf16 1.0 everywhere, so the swizzle is irrelevant)
  for (int i = threadIdx.x; ONES && i < NS * kAtomBytes / 16; i += kThreads) {
    const int st = i / (kAtomBytes / 16), o16 = i % (kAtomBytes / 16);
    *reinterpret_cast<uint4*>(sV + st * VTILE + NA * kAtomBytes + o16 * 16) =
        make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  }
#if LR_ATTN_STAGGER > 0
  if (!Cfg::kOnePerSm && warp == 3 && lane == 0) {   // warp 3 has no role in the one-tile configuration
    bool second;
    if (LR_ATTN_STAGGER_MODE == 1) {
      unsigned nsm;
      asm volatile("mov.u32 %0, %%nsmid;" : "=r"(nsm));
      const unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
      second = lin >= nsm && lin < 2 * nsm;
    } else {
#if LR_ATTN_STAGGER_MODE == 2
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      unsigned long long* slot = &g_attn_sm_arrivals[smid & 1023];
      unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(slot), seen;
      do {
        seen = old;
        const unsigned long long next = (seen >> 32) == epoch ? seen + 1 : ((unsigned long long)(epoch) << 32) | 1ull;
        old = atomicCAS(slot, seen, next);
      } while (old != seen);
      second = (seen >> 32) == epoch && (seen & 0xffffffffull) == 1ull;
#else
      second = false;
#endif
    }
    if (second) {
      const long long t0 = clock64();
      while (clock64() - t0 < LR_ATTN_STAGGER) {
      }
    }
  }
#endif
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // NT = 2: S_A 0, S_B 128, O_A 256, O_B 384;  NT = 1: S 0, O 128
  const uint32_t tm_S[2] = {tmem_base, tmem_base + 128};
  const uint32_t tm_O[2] = {tmem_base + (NT == 2 ? 256 : (PTMEM ? 192 : (PHYB ? 160 : 128))), tmem_base + 384};
  const uint32_t tm_P = tmem_base + 128;   // PTMEM: 64 columns of packed bf16 pairs

  // register re-allocation between warpgroups: the softmax threads keep a whole 128-column S row in registers
  if (warp < 4) {
  if constexpr (NT == 1) {
#if LR_ATTN_AUX_REGS == 40
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
#else
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LR_ATTN_AUX_REGS));
#endif
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  }
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int g = 0;  // K/V blocks of the earlier tiles of this CTA (ring positions and phases continue across tiles)
      for (int ti = 0; ti < n_tiles; ++ti) {
        set_tile(ti);
        if (n == 0) continue;
        // K and V rings advance independently (the MMA thread consumes them out of lock-step)
        int kj = 0, vj = 0;
        while (kj < n || vj < n) {
          [[maybe_unused]] const int kv_before = kj + vj;
          if (kj < n && mbar_test_wait(&k_empty[(g + kj) % NS], (((g + kj) / NS) & 1) ^ 1)) {
            if (kj == 0) {
              // Q of this tile. For a later tile the Q buffer is free once every S MMA of the previous tile - its only
              // readers - has retired: with a single-stage ring the wait above says exactly that; with two stages
              // the commit of the previous tile's LAST S MMA (block g - 1) is awaited explicitly.
              if (NS > 1 && g > 0) mbar_wait(&k_empty[(g - 1) % NS], ((g - 1) / NS) & 1);
              const uint32_t q_bytes = (nblk[0] > 0 ? TILE : 0) + (nblk[1] > 0 ? TILE : 0);
              mbar_arrive_expect_tx(q_full, q_bytes);
#pragma unroll
              for (int x = 0; x < 2; ++x)
                if (nblk[x] > 0)
                  for (int a = 0; a < NA; ++a)
                    tma_load_2d(sQ + x * TILE + a * kAtomBytes, &tm_qkv, q_full, q_col0 + head * HD + a * 32,
                                slot_row0 + m0 + x * 128);
            }
            const int s = (g + kj) % NS;
            if ((LR_ATTN_KO & 128) && kj > 0) {
              mbar_arrive(&k_full[s]);
            } else {
              mbar_arrive_expect_tx(&k_full[s], TILE);
              for (int a = 0; a < NA; ++a)
                tma_load_2d(sK + s * TILE + a * kAtomBytes, &tm_qkv, &k_full[s], k_col0 + kv_head * HD + a * 32,
                            slot_row0 + start + kj * 128);
              if (LR_ATTN_L2_PREFETCH && NS == 1 && kj + 1 < n)
                for (int a = 0; a < NA; ++a) {
                  tma_prefetch_l2_2d(&tm_qkv, k_col0 + kv_head * HD + a * 32, slot_row0 + start + (kj + 1) * 128);
                  tma_prefetch_l2_2d(&tm_qkv, v_col0 + kv_head * HD + a * 32, slot_row0 + start + (kj + 1) * 128);
                }
            }
            ++kj;
          }
          if (vj < n && vj < kj && mbar_test_wait(&v_empty[(g + vj) % NS], (((g + vj) / NS) & 1) ^ 1)) {
            const int s = (g + vj) % NS;
            if ((LR_ATTN_KO & 128) && vj > 0) {
              mbar_arrive(&v_full[s]);
            } else {
              mbar_arrive_expect_tx(&v_full[s], TILE);
              for (int a = 0; a < NA; ++a)
                tma_load_2d(sV + s * VTILE + a * kAtomBytes, &tm_qkv, &v_full[s], v_col0 + kv_head * HD + a * 32,
                            slot_row0 + start + vj * 128);
            }
            ++vj;
          }
          if (LR_ATTN_POLL_SLEEP > 0 && kj + vj == kv_before) poll_idle();
        }
        g += n;
      }
    }
    __syncwarp();
  } else if (warp == 1 || (warp == 3 && NT == 2)) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 -> tile A, warp 3 -> tile B
    // One thread per tile: a tcgen05.mma costs ~90 cycles of issue latency (measured with tools/attn_trace.py),
    // more than these 128x128x16 / 128x112x16 MMAs take to execute, so two issuing threads keep the tensor core fed.
    // Each tile advances as soon as ITS inputs are ready: S_x(j+1) the moment the softmax warps hold S_x(j) in
    // registers, O_x += P_x(j) V_j the moment P_x(j) is in smem. A K (V) stage goes back to the producer when both
    // threads have arrived on its `empty` barrier (tcgen05.commit after the last MMA that reads it, or a plain arrive
    // for blocks a tile does not need).
    if (kMmaWarp || lane == 0) {
      const int x = warp == 1 ? 0 : 1;
      int g = 0, tq = 0;  // K/V blocks / non-empty tiles already processed by this CTA
#if LR_ATTN_DESC32
      // descriptor low words of the four operand buffers, made opaque so that they stay in four registers instead of
      // being re-derived from the shared-memory base (shift + mask + or) in front of every tcgen05.mma
      uint32_t q_lo0 = umma_desc_lo(smem_u32(sQ + x * TILE), 16), k_lo0 = umma_desc_lo(smem_u32(sK), 16);
      uint32_t p_lo0 = umma_desc_lo(smem_u32(sP + x * kPBytes), 16), v_lo0 = umma_desc_lo(smem_u32(sV), kAtomBytes);
      asm volatile("" : "+r"(q_lo0), "+r"(k_lo0), "+r"(p_lo0), "+r"(v_lo0));
#endif
      for (int ti = 0; ti < n_tiles; ++ti) {
      set_tile(ti);
      if (n == 0) continue;
      const int nx = nblk[x];
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, HDX) | (1u << 16);  // B (= [V | 1]) is MN-major
      auto issue_s = [&](int ks) {  // S_x = Q_x K^T : A = Q (K-major, SW64), B = K stage ks (K-major, SW64)
        const uint32_t qa = smem_u32(sQ + x * TILE), ka = smem_u32(sK + ks * TILE);
#if LR_ATTN_HOIST_DESC
        const uint64_t qd = umma_desc(qa, 16, 512, kLayoutSW64), kd = umma_desc(ka, 16, 512, kLayoutSW64);
#endif
#if LR_ATTN_DESC32
        const uint32_t q_lo = q_lo0, k_lo = k_lo0 + ks * (TILE >> 4);
#endif
#pragma unroll
        for (int kk = 0; kk < ((LR_ATTN_KO & 32) ? 0 : HD / 16); ++kk) {
          const uint32_t off = (kk >> 1) * kAtomBytes + (kk & 1) * 32;
#if LR_ATTN_DESC32
          constexpr uint32_t hi = umma_desc_hi(512, kLayoutSW64);
          umma_bf16_ss(tm_S[x], desc_join(q_lo + (off >> 4), hi), desc_join(k_lo + (off >> 4), hi), idesc_s, kk != 0);
#elif LR_ATTN_HOIST_DESC
          // the start-address field holds (addr >> 4) in 14 bits; shared memory ends below 2^18, so the add cannot carry
          umma_bf16_ss(tm_S[x], qd + (off >> 4), kd + (off >> 4), idesc_s, kk != 0);
#else
          umma_bf16_ss(tm_S[x], umma_desc(qa + off, 16, 512, kLayoutSW64), umma_desc(ka + off, 16, 512, kLayoutSW64),
                       idesc_s, kk != 0);
#endif
        }
        umma_commit(&s_full[x]);
        umma_commit(&k_empty[ks]);
      };
      // kk0 .. kk1: the k-steps (16 keys each) to issue; the commits follow the last one (kk1 == 8)
      auto issue_pv = [&](int vs, bool acc, int kk0, int kk1) {  // O_x += P_x V : A = P (K-major, SW128), B = V (MN-major, SW64)
        if constexpr (PTMEM) {   // A = P from TMEM: MMA kk consumes keys [16 kk, 16 kk + 16) = 8 packed columns
          const uint32_t va = smem_u32(sV + vs * VTILE);
#if LR_ATTN_DESC32
          const uint32_t v_lo = v_lo0 + vs * (VTILE >> 4);
#endif
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if (kk < kk0 || kk >= kk1) continue;
#if LR_ATTN_DESC32
            umma_bf16_ts(tm_O[x], tm_P + kk * 8, desc_join(v_lo + ((kk * 1024) >> 4), umma_desc_hi(512, kLayoutSW64)),
                         idesc_o, acc || kk != 0);
#else
            umma_bf16_ts(tm_O[x], tm_P + kk * 8, umma_desc(va + kk * 1024, kAtomBytes, 512, kLayoutSW64), idesc_o,
                         acc || kk != 0);
#endif
          }
          if (kk1 == 8) {
            umma_commit(&pv_done[x]);
            umma_commit(&v_empty[vs]);
          }
          return;
        }
        const uint32_t pa = smem_u32(sP + x * kPBytes), va = smem_u32(sV + vs * VTILE);
#if LR_ATTN_HOIST_DESC
        const uint64_t pd = umma_desc(pa, 16, 1024, kLayoutSW128), vd = umma_desc(va, kAtomBytes, 512, kLayoutSW64);
#endif
#if LR_ATTN_DESC32
        const uint32_t p_lo = p_lo0, v_lo = v_lo0 + vs * (VTILE >> 4);
#endif
#pragma unroll
        for (int kk = 0; kk < ((LR_ATTN_KO & 16) ? 0 : 8); ++kk) {
          if (kk < kk0 || kk >= kk1) continue;
          if (PHYB && kk < 4) {   // keys 0..63: A = P from TMEM
#if LR_ATTN_DESC32
            umma_bf16_ts(tm_O[x], tm_P + kk * 8, desc_join(v_lo + ((kk * 1024) >> 4), umma_desc_hi(512, kLayoutSW64)),
                         idesc_o, acc || kk != 0);
#else
            umma_bf16_ts(tm_O[x], tm_P + kk * 8, umma_desc(va + kk * 1024, kAtomBytes, 512, kLayoutSW64), idesc_o,
                         acc || kk != 0);
#endif
            continue;
          }
#if LR_ATTN_DESC32
          umma_bf16_ss(tm_O[x], desc_join(p_lo + (((kk >> 2) * (kPBytes / 2) + (kk & 3) * 32) >> 4), umma_desc_hi(1024, kLayoutSW128)),
                       desc_join(v_lo + ((kk * 1024) >> 4), umma_desc_hi(512, kLayoutSW64)), idesc_o, acc || kk != 0);
#elif LR_ATTN_HOIST_DESC
          umma_bf16_ss(tm_O[x], pd + (((kk >> 2) * (kPBytes / 2) + (kk & 3) * 32) >> 4), vd + ((kk * 1024) >> 4), idesc_o,
                       acc || kk != 0);
#else
          umma_bf16_ss(tm_O[x], umma_desc(pa + (kk >> 2) * (kPBytes / 2) + (kk & 3) * 32, 16, 1024, kLayoutSW128),
                       umma_desc(va + kk * 1024, kAtomBytes, 512, kLayoutSW64), idesc_o, acc || kk != 0);
#endif
        }
        if (kk1 == 8) {
          umma_commit(&pv_done[x]);
          umma_commit(&v_empty[vs]);
        }
      };
      if (nx > 0) {
        mbar_wait(q_full, tq & 1);
        mbar_wait(&k_full[g % NS], (g / NS) & 1);
        if (g > 0) mbar_wait(&s_free[x], (g - 1) & 1);  // the S row of the previous tile's last block has been read
        if (NT == 2 && x == 1 && nblk[0] > 0) {
          // Start tile B half a softmax period after tile A: both softmax warpgroups share the SM's 16 ex2/clk, and
          // the exponential phase is about half of an iteration, so in anti-phase each runs its exponentials alone.
          mbar_wait(&s_free[0], 0);
          const long long t0 = clock64();
          while (clock64() - t0 < kStaggerCycles) {
          }
        }
        tc_fence_after();
        LR_MMA_LEAD_BEGIN
        issue_s(g % NS);
        LR_MMA_LEAD_END
        int s_next = 1, pv_next = 0;
        [[maybe_unused]] bool half_issued = false;
        while (pv_next < nx) {
          [[maybe_unused]] const int ev_before = s_next + pv_next + int(half_issued);
          if (s_next < nx && LR_MMA_VOTE(mbar_test_wait(&s_free[x], (g + s_next - 1) & 1) &&
                                         mbar_test_wait(&k_full[(g + s_next) % NS], ((g + s_next) / NS) & 1))) {
            tc_fence_after();
            ATTN_TRACE(0, x * 4 + 0, s_next - 1);
            LR_MMA_LEAD_BEGIN
            issue_s((g + s_next) % NS);
            LR_MMA_LEAD_END
            ++s_next;
          }
          bool pv_ready;
          if constexpr (LR_ATTN_P_HALF && SPLIT == 1) {
            if (!half_issued && LR_MMA_VOTE(mbar_test_wait(&p_half[x], (g + pv_next) & 1) &&
                                            mbar_test_wait(&v_full[(g + pv_next) % NS], ((g + pv_next) / NS) & 1))) {
              tc_fence_after();
              LR_MMA_LEAD_BEGIN
              issue_pv((g + pv_next) % NS, pv_next > 0, 0, 4);   // keys 0..63 of P_x(j) are stored (and O_x rescaled)
              LR_MMA_LEAD_END
              half_issued = true;
            }
            pv_ready = half_issued && LR_MMA_VOTE(mbar_test_wait(&p_full[x], (g + pv_next) & 1));
          } else {
            pv_ready = LR_MMA_VOTE(mbar_test_wait(&p_full[x], (g + pv_next) & 1) &&
                                   mbar_test_wait(&v_full[(g + pv_next) % NS], ((g + pv_next) / NS) & 1));
          }
          if (pv_ready) {
            tc_fence_after();
            ATTN_TRACE(0, x * 4 + 1, pv_next);
            LR_MMA_LEAD_BEGIN
            if constexpr (LR_ATTN_P_HALF && SPLIT == 1) {
              issue_pv((g + pv_next) % NS, true, 4, 8);
            } else {
              issue_pv((g + pv_next) % NS, pv_next > 0, 0, 8);
            }
            if (pv_next + 1 == nx) umma_commit(&o_final[x]);
            LR_MMA_LEAD_END
            half_issued = false;
            ATTN_TRACE(0, x * 4 + 2, pv_next);
            ++pv_next;
          }
          if (LR_ATTN_POLL_SLEEP > 0 && s_next + pv_next + int(half_issued) == ev_before) poll_idle();
        }
      }
      for (int j = nx; j < n; ++j) {  // K/V blocks only the other tile reads (NT == 2: one tile pair per CTA, g == 0)
        const int st = j % NS;
        if (j >= NS) {
          mbar_wait(&k_empty[st], ((j - NS) / NS) & 1);
          mbar_wait(&v_empty[st], ((j - NS) / NS) & 1);
        }
        LR_MMA_LEAD_BEGIN
        mbar_arrive(&k_empty[st]);
        mbar_arrive(&v_empty[st]);
        LR_MMA_LEAD_END
      }
      g += n;
      ++tq;
      }
    }
    __syncwarp();
  }
  } else {
    if constexpr (NT == 1 && SPLIT == 2) {
      // experiment (tools/attn_variants.py --split): two threads per row, 384 threads x 80 registers at launch,
      // re-split 128 x 40 + 256 x 96
#ifndef LR_ATTN_SPLIT_REGS
#define LR_ATTN_SPLIT_REGS 96
#endif
      asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(LR_ATTN_SPLIT_REGS));
    } else if constexpr (NT == 1) {
      // 2 CTAs/SM: the CTA's 256 x 128 registers are re-split 128 x AUX + 128 x (256 - AUX) (40 + 208 leaves 8 unused)
      asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(LR_ATTN_AUX_REGS <= 48 ? 208 : 256 - LR_ATTN_AUX_REGS));
    } else if constexpr (SPLIT == 1) {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    }
    // ------------------------------------------------------------------ softmax warpgroups
    constexpr int NCH = 4 / SPLIT;                 // 32-column chunks of S owned by one thread
    const int sw = warp - 4;
    const int x = sw / (4 * SPLIT);                // tile A or B
    const int h = (sw >> 2) % SPLIT;               // column half (SPLIT == 2)
    const int q = warp & 3;                        // TMEM lane quarter
    const int r = q * 32 + lane;                   // row inside the tile
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    uint8_t* prow = sP + x * kPBytes + r * 128;
    const bool tr = (q == 0 && lane == 0 && h == 0);
    int g = 0, tq = 0;  // K/V blocks / non-empty tiles already processed by this CTA (barrier phases)
    for (int ti = 0; ti < n_tiles; ++ti) {
    set_tile(ti);
    const int row_abs = m0 + x * 128 + r;
    float m_ref = -INFINITY;
    float l_reg[4] = {0.f, 0.f, 0.f, 0.f};  // row sum kept in registers when there are no ones columns (!ONES)
    const int nx = nblk[x];
    int my_lo = 0, my_hi = 0;  // segment layout: this row's key range (absolute rows)
    if (seg && row_abs < rows_per_seq) {
      my_lo = row_lo[row_abs];
      my_hi = row_hi[row_abs];
    }
    uint32_t sv[NCH][32];          // the S row of the current block (LR_ATTN_PRELOAD: refilled with the next block's)
    [[maybe_unused]] bool have = false;   // sv already holds S(j), its row max is mx_pre and s_free(j) has been signalled
    [[maybe_unused]] float mx_pre = 0.f;
    for (int j = 0; j < nx; ++j) {
      const bool preloaded = LR_ATTN_PRELOAD && have;
      have = false;
      if (!preloaded) {
        if (tr) ATTN_TRACE(1 + x, 0, j);
        softmax_wait(&s_full[x], (g + j) & 1);
        tc_fence_after();
        if (tr) ATTN_TRACE(1 + x, 1, j);
      }
      const int kv0 = start + j * 128;
      const bool need_mask = seg || (kv0 + 128 > kv_end[x]) || (CAUSAL && kv0 + 127 > m0 + x * 128);
      // whole S row -> registers, then give the TMEM buffer back to the MMA warp
      float mxc[8];  // 8 independent chains instead of one long dependent FMNMX chain
#pragma unroll
      for (int c = 0; c < 8; ++c) mxc[c] = -INFINITY;
      [[maybe_unused]] int nvalid = 128, vmin = 128, vmax = 128;   // leading valid columns: this row / warp min / warp max
      if (LR_ATTN_CHUNK_MASK && need_mask && !seg) {
        int lim = kv_end[x] - kv0;
        if (CAUSAL) lim = min(lim, row_abs + 1 - kv0);
        nvalid = max(0, min(128, lim));
        vmin = __reduce_min_sync(0xffffffffu, nvalid);
        vmax = __reduce_max_sync(0xffffffffu, nvalid);
      }
      auto chunk_dead = [&](int c) { return LR_ATTN_CHUNK_MASK != 0 && (h * NCH + c) * 32 >= vmax; };  // warp-uniform
      auto mask_chunk = [&](int c) {
        if (LR_ATTN_CHUNK_MASK && !seg) {
          const int base = (h * NCH + c) * 32;
          // every lane keeps the whole chunk, or no lane keeps anything: a dead chunk is never read again (no max, no
          // exponentials). Writing -inf into it instead costs 1.2 KB of spills (ptxas), so it is simply left alone.
          if (base + 32 <= vmin || base >= vmax) return;
#pragma unroll
          for (int i = 0; i < 32; ++i) sv[c][i] = (base + i < nvalid) ? sv[c][i] : 0xff800000u;
          return;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = kv0 + (h * NCH + c) * 32 + i;
          const bool ok = seg ? (col >= my_lo && col < my_hi) : (col < kv_end[x] && (!CAUSAL || col <= row_abs));
          sv[c][i] = ok ? sv[c][i] : 0xff800000u;  // -inf
        }
      };
      auto max_chunk = [&](int c) {
#if LR_ATTN_MAX3
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float& m = mxc[(c * 4 + (i & 3)) & 7];
          m = fmax3(m, __uint_as_float(sv[c][2 * i]), __uint_as_float(sv[c][2 * i + 1]));
        }
#else
#pragma unroll
        for (int i = 0; i < 32; ++i) mxc[(c * 2 + (i >> 4)) & 7] = fmaxf(mxc[(c * 2 + (i >> 4)) & 7], __uint_as_float(sv[c][i]));
#endif
      };
      float mx = 0.f;
      // speculative reference maximum (see LR_ATTN_SPEC_MAX): CTA-uniform
      const bool spec = LR_ATTN_SPEC_MAX != 0 && SPLIT == 1 && !LR_ATTN_PIPE_LD && !LR_ATTN_PRELOAD && j > 0 && !need_mask;
      if (!preloaded) {
      if constexpr (SPLIT == 1 && LR_ATTN_PIPE_LD) {
        // chunk c+1 is in flight while chunk c is masked and reduced
        tmem_ld_32x32(tm_S[x] + lane_addr, sv[0]);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          tmem_ld_wait_dep(sv[c]);
          if (c + 1 < NCH) tmem_ld_32x32(tm_S[x] + lane_addr + (c + 1) * 32, sv[c + 1]);
          if (need_mask) mask_chunk(c);
          if (!(LR_ATTN_KO & 4) && !chunk_dead(c)) max_chunk(c);
        }
        if (tr) ATTN_TRACE(1 + x, 2, j);
      } else {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          tmem_ld_32x32(tm_S[x] + lane_addr + (h * NCH + c) * 32, sv[c]);
        tmem_ld_wait();
        if constexpr (SPLIT == 1 && LR_ATTN_EARLY_SFREE) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[x]);
        }
        if (tr) ATTN_TRACE(1 + x, 2, j);
        if (need_mask) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            mask_chunk(c);
            if (!(LR_ATTN_KO & 4) && !chunk_dead(c)) max_chunk(c);
          }
        } else if (!spec) {
#pragma unroll
          for (int c = 0; c < NCH; ++c)
            if (!(LR_ATTN_KO & 4)) max_chunk(c);
        }
      }
      if (!spec) mx = fmaxf(fmaxf(fmaxf(mxc[0], mxc[1]), fmaxf(mxc[2], mxc[3])), fmaxf(fmaxf(mxc[4], mxc[5]), fmaxf(mxc[6], mxc[7])));
      if constexpr (SPLIT == 2) {
        // The other 64 columns of this row live in a thread of the partner warpgroup: exchange the partial maxima
        // through smem. The S buffer is released only AFTER the exchange, so S_x(j+1) - and with it the next write
        // to this buffer - cannot happen before every thread has read its partner's value of block j.
        float* mb = mbuf + x * 256;
        mb[h * 128 + r] = mx;
        asm volatile("bar.sync %0, 256;" ::"r"(1 + x) : "memory");
        mx = fmaxf(mx, mb[(h ^ 1) * 128 + r]);
      }
      if constexpr (!(SPLIT == 1 && LR_ATTN_EARLY_SFREE)) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[x]);
      }
      } else {
        mx = mx_pre;   // S(j) was read and reduced under the exponentials of block j-1
      }
      // lazy rescale: move the reference max only when it grew by more than the threshold
      float alpha = 1.f, msafe = (m_ref == -INFINITY) ? 0.f : m_ref;
      bool grow = false;
      auto update_ref = [&](float mxs) {   // mxs = row max x scale (scale > 0, so max commutes with the scaling)
        grow = mxs > m_ref + kRescaleThreshold || (m_ref == -INFINITY && mxs > -INFINITY);
        if (grow) {
          alpha = exp2f(m_ref - mxs);  // 0 when m_ref = -inf
          m_ref = mxs;
        }
        msafe = (m_ref == -INFINITY) ? 0.f : m_ref;
      };
      if (!spec) {
        update_ref(mx * scale_log2);
        if constexpr (!ONES) {
          if (grow) {
#pragma unroll
            for (int t = 0; t < 4; ++t) l_reg[t] *= alpha;
          }
        }
      }
      if (tr) ATTN_TRACE(1 + x, 3, j);
      // P = exp2(S*scale - m_ref) (masked entries: exp2(-inf) = 0) -> bf16, 32 columns (one chunk) at a time
      // MASKED (compile-time tag): the instantiation for blocks that need masking consults chunk_dead(); the one for
      // all other blocks stays a single straight-line region (a warp-uniform branch per chunk keeps the scheduler from
      // interleaving the FFMA / MUFU / F2FP streams of neighbouring chunks: measured +3 % when it sat in the common path)
      auto exp_chunk = [&](int c, uint32_t (&pk)[16], auto masked_tag) {
        if (decltype(masked_tag)::value && chunk_dead(c)) {   // masked for every row of this warp: P = 0, no exponentials
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
          return;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
#if LR_ATTN_FFMA2
          float x0, x1;
          ffma2(x0, x1, sv[c][2 * i], sv[c][2 * i + 1], scale_log2, -msafe);
          const float p0 = softmax_exp2(x0, 2 * i), p1 = softmax_exp2(x1, 2 * i + 1);
#else
          const float p0 = softmax_exp2(fmaf(__uint_as_float(sv[c][2 * i]), scale_log2, -msafe), 2 * i);
          const float p1 = softmax_exp2(fmaf(__uint_as_float(sv[c][2 * i + 1]), scale_log2, -msafe), 2 * i + 1);
#endif
          if constexpr (!ONES) {
            l_reg[(2 * i) & 3] += p0;
            l_reg[(2 * i + 1) & 3] += p1;
          }
          pk[i] = pack_bf16x2(p0, p1);
        }
      };
      // -> smem (or this thread's TMEM lane), so the stores issue under the MUFU-bound exponentials of the next chunk.
      // 128 columns = 2 atoms x 8 chunks of 16 B; chunk position XOR (row & 7) = the 128B swizzle.
      auto store_chunk = [&](int c, const uint32_t (&pk)[16]) {
        const int cg = h * NCH + c;  // chunk index inside the 128-column row
        if (PTMEM || (PHYB && cg < 2)) {   // 32 keys = 16 packed columns of this thread's own TMEM lane
          tmem_st_32x16(tm_P + lane_addr + cg * 16, pk);
          return;
        }
        uint8_t* base = prow + (cg >> 1) * (kPBytes / 2);
#if LR_ATTN_KO & 2
        {  // keep every exponential and pack alive (XOR of all packed words), store (practically) never
          uint32_t live = 0;
#pragma unroll
          for (int i = 0; i < 16; ++i) live ^= pk[i];
          if (live == 0x12345678u) *reinterpret_cast<uint4*>(base) = make_uint4(pk[0], pk[5], pk[9], live);
        }
#else
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int chunk = ((cg & 1) * 4 + t) ^ (r & 7);
          *reinterpret_cast<uint4*>(base + (chunk << 4)) = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
        }
#endif
      };
      // LR_ATTN_PRELOAD: may the next block be pulled in under this block's exponentials?
      [[maybe_unused]] bool next_ready = false;
      [[maybe_unused]] const bool next_ok = LR_ATTN_PRELOAD && SPLIT == 1 && !LR_ATTN_PIPE_LD && j + 1 < nx && !seg &&
                                            !(kv0 + 256 > kv_end[x]) && !(CAUSAL && kv0 + 128 + 127 > m0 + x * 128);
      auto preload_hook = [&](auto c_tag) {   // chunk c of S(j) has been consumed: its registers can take S(j+1)
        constexpr int c = decltype(c_tag)::value;
        if constexpr (LR_ATTN_PRELOAD != 0 && SPLIT == 1) {
          if (!next_ok) return;
          if (c <= 1 && !next_ready && mbar_test_wait(&s_full[x], (g + j + 1) & 1)) {
            next_ready = true;
            tc_fence_after();
            if (c == 1) tmem_ld_32x32(tm_S[x] + lane_addr, sv[0]);
          }
          if (next_ready) tmem_ld_32x32(tm_S[x] + lane_addr + c * 32, sv[c]);
        }
      };
      auto p_phase = [&](auto masked_tag) {
      [[maybe_unused]] uint32_t pk_first[16];
      if constexpr (LR_ATTN_EXP_FIRST != 0) exp_chunk(0, pk_first, masked_tag);   // registers only: runs under the wait below
      // P_x buffer and O_x may be touched only after O_x += P_x(j-1) V_(j-1) has retired
      if (j > 0) {
        softmax_wait(&pv_done[x], (g + j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {
#pragma unroll 1
          for (int c = h; c < HD / 32 + (ONES ? 1 : 0); c += SPLIT) {  // +1: the chunk that holds the row-sum column
            uint32_t ov[32];
            tmem_ld_32x32(tm_O[x] + lane_addr + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x32(tm_O[x] + lane_addr + c * 32, ov);
          }
          tmem_st_wait();
        }
      }
      if (tr) ATTN_TRACE(1 + x, 4, j);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (LR_ATTN_EXP_FIRST != 0 && c == 0) {
          store_chunk(0, pk_first);
        } else {
          uint32_t pk[16];
          exp_chunk(c, pk, masked_tag);
          store_chunk(c, pk);
        }
        if constexpr (LR_ATTN_PRELOAD != 0 && SPLIT == 1) {
          if (c == 0) preload_hook(std::integral_constant<int, 0>{});
          if (c == 1) preload_hook(std::integral_constant<int, 1>{});
          if (c == 2) preload_hook(std::integral_constant<int, 2>{});
          if (c == 3) preload_hook(std::integral_constant<int, 3>{});
        }
        if (LR_ATTN_P_HALF && SPLIT == 1 && c == NCH / 2 - 1) {   // keys 0..63 are stored: the MMA thread may issue the first half of P.V
          if constexpr (PTMEM || PHYB) tmem_st_wait();
          else fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_half[x]);
        }
      }
      };
      auto p_phase_spec = [&]() {
        [[maybe_unused]] float l_save[4] = {l_reg[0], l_reg[1], l_reg[2], l_reg[3]};
        uint32_t pk0[16], pk1[16];
        exp_chunk(0, pk0, std::false_type{});          // with the old reference
#pragma unroll
        for (int c = 0; c < NCH; ++c)                  // independent of the exponentials: fills their idle issue slots
          if (!(LR_ATTN_KO & 4)) max_chunk(c);
        softmax_wait(&pv_done[x], (g + j - 1) & 1);    // spec implies j > 0
        tc_fence_after();
        exp_chunk(1, pk1, std::false_type{});
        const float mxs = fmaxf(fmaxf(fmaxf(mxc[0], mxc[1]), fmaxf(mxc[2], mxc[3])),
                                fmaxf(fmaxf(mxc[4], mxc[5]), fmaxf(mxc[6], mxc[7]))) * scale_log2;
        update_ref(mxs);
        if (__any_sync(0xffffffffu, grow)) {
          // some row of this warp outgrew the threshold: rescale O and redo the first 64 keys with the new reference
          // (rows that did not grow have alpha = 1 and the same reference: they reproduce their values)
          if constexpr (!ONES) {
#pragma unroll
            for (int t = 0; t < 4; ++t) l_reg[t] = l_save[t] * alpha;
          }
#pragma unroll 1
          for (int c = 0; c < HD / 32 + (ONES ? 1 : 0); ++c) {
            uint32_t ov[32];
            tmem_ld_32x32(tm_O[x] + lane_addr + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x32(tm_O[x] + lane_addr + c * 32, ov);
          }
          tmem_st_wait();
          exp_chunk(0, pk0, std::false_type{});
          exp_chunk(1, pk1, std::false_type{});
        }
        if (tr) ATTN_TRACE(1 + x, 4, j);
        store_chunk(0, pk0);
        store_chunk(1, pk1);
        if (LR_ATTN_P_HALF) {   // keys 0..63 are stored: the MMA thread may issue the first half of P.V
          if constexpr (PTMEM || PHYB) tmem_st_wait();
          else fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_half[x]);
        }
#pragma unroll
        for (int c = 2; c < NCH; ++c) {
          uint32_t pk[16];
          exp_chunk(c, pk, std::false_type{});
          store_chunk(c, pk);
        }
      };
      if (LR_ATTN_CHUNK_MASK != 0 && need_mask && !seg) p_phase(std::true_type{});
      else if (spec) p_phase_spec();
      else p_phase(std::false_type{});
#if !(LR_ATTN_KO & 2) && !(LR_ATTN_KO & 64)
      if constexpr (PTMEM || PHYB) tmem_st_wait();   // the tcgen05.st of the P row have landed
      if constexpr (!PTMEM) fence_proxy_async_smem();  // P (generic-proxy stores) must be visible to the tensor core's async proxy
#endif
      if (tr) ATTN_TRACE(1 + x, 5, j);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[x]);
      if (tr) ATTN_TRACE(1 + x, 6, j);
      if constexpr (LR_ATTN_PRELOAD != 0 && SPLIT == 1) {
        if (next_ready) {   // S(j+1) is (arriving) in sv: finish the read, reduce, hand the S buffer back
          tmem_ld_wait();
          float m8[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) m8[c] = -INFINITY;
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) m8[(c * 2 + (i >> 4)) & 7] = fmaxf(m8[(c * 2 + (i >> 4)) & 7], __uint_as_float(sv[c][i]));
          mx_pre = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[x]);
          have = true;
        }
      }
    }
    // epilogue: O / l -> bf16 -> global; rows outside the valid run are zero-filled
    const bool row_in_slot = row_abs < (seq_base ? end : rows_per_seq);  // packed: the next rows are another sequence
    const bool row_valid = row_abs >= lo[x] && row_abs < hi[x];
    bf16* orow = o + size_t(slot_row0 + row_abs) * ld_o + head * HD;
    if (nx > 0) {
      if (tr) ATTN_TRACE(1 + x, 7, 0);   // last p_full arrived: the drain of the tile starts
      mbar_wait(&o_final[x], tq & 1);
      // the last block's pv_done phase completed together with o_final (same commit point); observing it here keeps
      // every phase of that barrier waited-on before the next tile arrives on it again (compute-sanitizer synccheck)
      mbar_wait(&pv_done[x], (g + nx - 1) & 1);
      tc_fence_after();
      if (tr) ATTN_TRACE(1 + x, 7, 1);   // last P.V retired
      float inv;
      if constexpr (!ONES) {
        const float l_sum = (l_reg[0] + l_reg[1]) + (l_reg[2] + l_reg[3]);
        inv = l_sum > 0.f ? 1.f / l_sum : 0.f;
      } else {
        uint32_t lv[32];
        tmem_ld_32x32(tm_O[x] + lane_addr + HD, lv);  // column HD = sum_j P_ij (accumulated by the PV MMAs)
        tmem_ld_wait();
        const float l_sum = __uint_as_float(lv[0]);
        inv = l_sum > 0.f ? 1.f / l_sum : 0.f;
      }
      if constexpr (LR_ATTN_EPI_STAGE && SPLIT == 1) {
        constexpr int CH = HD / 8;               // 16-byte pieces per output row
        constexpr int PITCH = Cfg::kStagePitch;
        auto stage_row = [&](int rr) -> uint8_t* {   // row rr (0..31) of this warp
          if constexpr (PTMEM) return sStage + (q * 32 + rr) * PITCH;
          // this warp's own rows of the retired P buffer: 4 KB in each of the two 64-column atoms, 16 rows in each
          else return sP + x * kPBytes + (rr >> 4) * (kPBytes / 2) + q * 4096 + (rr & 15) * PITCH;
        };
        uint8_t* my_row = stage_row(lane);
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t ov[32];
          tmem_ld_32x32(tm_O[x] + lane_addr + c * 32, ov);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(ov[8 * t]) * inv, __uint_as_float(ov[8 * t + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(ov[8 * t + 2]) * inv, __uint_as_float(ov[8 * t + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(ov[8 * t + 4]) * inv, __uint_as_float(ov[8 * t + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(ov[8 * t + 6]) * inv, __uint_as_float(ov[8 * t + 7]) * inv);
            const int piece = c * 4 + t;
            *reinterpret_cast<uint4*>(my_row + ((HD == 128 ? (piece ^ (lane & 7)) : piece) << 4)) = u;
          }
        }
        __syncwarp();
        const int row0 = m0 + x * 128 + q * 32;   // first row of this warp
        const int slot_end = seq_base ? end : rows_per_seq;
#pragma unroll 1
        for (int it = 0; it < CH; ++it) {
          const int idx = it * 32 + lane, rr = idx / CH, piece = idx - rr * CH;
          const int ra = row0 + rr;
          if (ra < slot_end) {
            uint4 u = make_uint4(0, 0, 0, 0);
            if (ra >= lo[x] && ra < hi[x])
              u = *reinterpret_cast<const uint4*>(stage_row(rr) + ((HD == 128 ? (piece ^ (rr & 7)) : piece) << 4));
            *reinterpret_cast<uint4*>(o + size_t(slot_row0 + ra) * ld_o + head * HD + piece * 8) = u;
          }
        }
        __syncwarp();   // the staging rows are this warp's P rows of the next tile
      } else {
#pragma unroll 1
      for (int c = h; c < HD / 32; c += SPLIT) {
        uint32_t ov[32];
        tmem_ld_32x32(tm_O[x] + lane_addr + c * 32, ov);
        tmem_ld_wait();
        if (row_in_slot) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            uint4 u = make_uint4(0, 0, 0, 0);
            if (row_valid) {
              u.x = pack_bf16x2(__uint_as_float(ov[8 * t]) * inv, __uint_as_float(ov[8 * t + 1]) * inv);
              u.y = pack_bf16x2(__uint_as_float(ov[8 * t + 2]) * inv, __uint_as_float(ov[8 * t + 3]) * inv);
              u.z = pack_bf16x2(__uint_as_float(ov[8 * t + 4]) * inv, __uint_as_float(ov[8 * t + 5]) * inv);
              u.w = pack_bf16x2(__uint_as_float(ov[8 * t + 6]) * inv, __uint_as_float(ov[8 * t + 7]) * inv);
            }
            *reinterpret_cast<uint4*>(orow + c * 32 + t * 8) = u;
          }
        }
      }
      }
    } else if (row_in_slot && h == 0) {
#pragma unroll 1
      for (int c = 0; c < HD / 8; ++c) *reinterpret_cast<uint4*>(orow + c * 8) = make_uint4(0, 0, 0, 0);
    }
    if (tr) ATTN_TRACE(1 + x, 7, 2);     // O written to global memory
    g += n;
    if (n > 0) ++tq;
    // the O accumulator and the P buffer of this tile are free again: the next tile's first P.V is issued only after
    // these threads have arrived on p_full for it, i.e. after the tcgen05.ld of the epilogue above have completed
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
#ifdef LR_ATTN_TRACE
  if (threadIdx.x == 0 && g_attn_cta) {
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    const long long id = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
    g_attn_cta[id * 4 + 0] = smid;
    g_attn_cta[id * 4 + 1] = cta_t0;
    g_attn_cta[id * 4 + 2] = attn_globaltimer();
    g_attn_cta[id * 4 + 3] = nblk[0] + (NT == 2 ? nblk[1] : 0);
  }
#endif
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn attn_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    (void)cudaGetLastError();
  }
  return fn;
}

static std::atomic<unsigned> g_launch_epoch{0};   // tags the per-SM arrival counters of one launch (LR_ATTN_STAGGER_MODE 2 experiment)

template <int HD, bool CAUSAL, int SPLIT, int NT>
static int launch_attn_tc(const void* base, int total_rows, int ld_qkv, int q_col0, int k_col0, int v_col0, void* o,
                          int ld_o, int n_seq, int rows_per_seq, const int* seq_base, const int* seq_start,
                          const int* seq_len, int n_heads, int kv_group, float scale, cudaStream_t stream,
                          const int* row_lo = nullptr, const int* row_hi = nullptr, bool multi_tile = false) {
  using Cfg = AttnTcCfg<HD, NT>;
  EncodeTiledFn fn = attn_encode_fn();
  if (!fn) return LR_ERR_NO_DRIVER;
  CUtensorMap tm;
  if (!cached_tmap_bf16(fn, &tm, base, total_rows, ld_qkv, ld_qkv, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B)) return LR_ERR_BAD_ARG;
  auto kern = attn_tc_kernel<HD, CAUSAL, SPLIT, NT>;
  {  // per-device attribute, set once per (kernel, device)
    static bool attr_done[64] = {};   // one array per instantiation of this launch template = per kernel
    cudaError_t e = ensure_smem_attr(kern, Cfg::kSmemBytes + ((SPLIT == 2 && NT == 1) ? 1024 : 0), attr_done);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  // several query tiles per CTA (see the kernel): single-stage K/V ring only, not for the segment layout
  const int nt = (rows_per_seq + 128 * NT - 1) / (128 * NT);
  int tile_mode = 0, gx = nt;
  // Measured (B200, profiles/r01_attention_multitile.txt): CLIP (5 tiles of 5 K/V blocks) 1.39 -> 1.29 ms, causal
  // decoder pairs 1.264 -> 1.252 ms; a 4900-token non-causal sequence (39 blocks per tile, fixed cost already
  // amortised, fewer and longer CTAs -> wave quantisation) 1.45 -> 1.60 ms, so long non-causal sequences keep one
  // tile per CTA.
  if (multi_tile && NT == 1 && !Cfg::kOnePerSm && row_lo == nullptr && nt > 1) {
    if (CAUSAL) {
      tile_mode = 1;                 // pairs {nt-1-x, x}
      gx = (nt + 1) / 2;
    } else if (nt <= 6) {
      tile_mode = nt;                // all tiles of a short sequence (a whole 577-token CLIP crop) in one CTA
      gx = 1;
    }
  }
  dim3 grid(gx, n_heads, n_seq);
  constexpr int kSmem = Cfg::kSmemBytes + ((SPLIT == 2 && NT == 1) ? 1024 : 0);   // row-max exchange of the experiment
  kern<<<grid, 128 + 128 * NT * SPLIT, kSmem, stream>>>(tm, reinterpret_cast<bf16*>(o), ld_o, rows_per_seq, seq_base,
                                                      seq_start, seq_len, q_col0, k_col0, v_col0, kv_group,
                                                      scale * 1.4426950408889634f, row_lo, row_hi, tile_mode, g_launch_epoch.fetch_add(1u, std::memory_order_relaxed) + 1u);
  return lr_launch_status();
}

// q, k, v must be column offsets into ONE row-major buffer (the fused qkv projection): base = q.
// seq_base == NULL: slot layout (total_rows = n_seq * rows_per_seq); else packed sequences (rows_per_seq = max length).
int attention_tc(const void* q, const void* k, const void* v, void* o, int ld_qkv, int ld_o, int total_rows, int n_seq,
                 int rows_per_seq, const int* seq_base, const int* seq_start, const int* seq_len, int n_heads,
                 int n_kv_heads, int head_dim, int causal, float scale, int split, cudaStream_t s) {
  const ptrdiff_t kd = (reinterpret_cast<const char*>(k) - reinterpret_cast<const char*>(q)) / 2;
  const ptrdiff_t vd = (reinterpret_cast<const char*>(v) - reinterpret_cast<const char*>(q)) / 2;
  if (n_kv_heads <= 0 || n_heads % n_kv_heads) return LR_ERR_BAD_ARG;
  const int kvw = n_kv_heads * head_dim, g = n_heads / n_kv_heads;
  if (kd < 0 || vd < 0 || kd + kvw > ld_qkv || vd + kvw > ld_qkv || n_heads * head_dim > ld_qkv) return LR_ERR_BAD_ARG;
#define LR_ATTN_ARGS q, total_rows, ld_qkv, 0, int(kd), int(vd), o, ld_o, n_seq, rows_per_seq, seq_base, seq_start, seq_len, n_heads, g, scale, s
#define LR_ATTN_CASE(HD_, CAUSAL_)                                                     \
  if (head_dim == HD_ && bool(causal) == CAUSAL_) {                                    \
    if (split == 2) return launch_attn_tc<HD_, CAUSAL_, 2, 2>(LR_ATTN_ARGS);           \
    if (split == 3) return launch_attn_tc<HD_, CAUSAL_, 1, 1>(LR_ATTN_ARGS);           \
    if (split == 4) return launch_attn_tc<HD_, CAUSAL_, 1, 1>(LR_ATTN_ARGS, nullptr, nullptr, true); \
    if (split == 6) {  /* two softmax threads per row, one tile per CTA: not with P in TMEM (row sums per thread) */ \
      if (AttnTcCfg<HD_, 1>::kPTmem || AttnTcCfg<HD_, 1>::kPHybrid) return LR_ERR_BAD_ARG;                            \
      return launch_attn_tc<HD_, CAUSAL_, 2, 1>(LR_ATTN_ARGS, nullptr, nullptr, true); \
    }                                                                                  \
    return launch_attn_tc<HD_, CAUSAL_, 1, 2>(LR_ATTN_ARGS);                           \
  }
  LR_ATTN_CASE(64, false)
  LR_ATTN_CASE(96, true)
#undef LR_ATTN_CASE
  if (head_dim == 96 && !causal) {  // Qwen2.5-VL vision tower (head_dim 80 zero-padded to 96): one tile per CTA only
    if (split == 2 || split == 1) return LR_ERR_BAD_ARG;
    return launch_attn_tc<96, false, 1, 1>(LR_ATTN_ARGS, nullptr, nullptr, split == 4);
  }
  if (head_dim == 128 && causal) {  // one CTA per SM either way (smem / TMEM, see AttnTcCfg); no split-softmax form
    if (split == 2) return LR_ERR_BAD_ARG;
    if (split == 1)  // two query tiles per CTA, row sums in registers
      return launch_attn_tc<128, true, 1, 2>(LR_ATTN_ARGS);
    return launch_attn_tc<128, true, 1, 1>(LR_ATTN_ARGS);
  }
#undef LR_ATTN_ARGS
  return LR_ERR_BAD_ARG;
}

// Segment layout (see the kernel): non-causal, head_dim 64 or 96, one query tile per CTA.
int attention_tc_seg(const void* q, const void* k, const void* v, void* o, int ld_qkv, int ld_o, int total_rows,
                     const int* row_lo, const int* row_hi, int n_heads, int head_dim, float scale, cudaStream_t s) {
  const ptrdiff_t kd = (reinterpret_cast<const char*>(k) - reinterpret_cast<const char*>(q)) / 2;
  const ptrdiff_t vd = (reinterpret_cast<const char*>(v) - reinterpret_cast<const char*>(q)) / 2;
  const int w = n_heads * head_dim;
  if (kd < 0 || vd < 0 || kd + w > ld_qkv || vd + w > ld_qkv) return LR_ERR_BAD_ARG;
  if (head_dim == 96)
    return launch_attn_tc<96, false, 1, 1>(q, total_rows, ld_qkv, 0, int(kd), int(vd), o, ld_o, 1, total_rows, nullptr,
                                           nullptr, nullptr, n_heads, 1, scale, s, row_lo, row_hi);
  if (head_dim == 64)
    return launch_attn_tc<64, false, 1, 1>(q, total_rows, ld_qkv, 0, int(kd), int(vd), o, ld_o, 1, total_rows, nullptr,
                                           nullptr, nullptr, n_heads, 1, scale, s, row_lo, row_hi);
  return LR_ERR_BAD_ARG;
}

}  // namespace lr

#ifdef LR_ATTN_TRACE
extern "C" int lr_attn_trace_set(long long* buf) {
  return static_cast<int>(cudaMemcpyToSymbol(lr::g_attn_trace, &buf, sizeof(buf)));
}
extern "C" int lr_attn_cta_set(long long* buf) {
  return static_cast<int>(cudaMemcpyToSymbol(lr::g_attn_cta, &buf, sizeof(buf)));
}
#endif

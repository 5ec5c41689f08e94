// recon.cu — kernels' entry points and the extern "C" ABI of include/dryv_recon.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
//             recon.cu recon_tables.cpp -o libdryv_recon.so        (see __graft_entry__.build)
// There is no CPU implementation behind this ABI: without a usable CUDA device every reconstructing
// entry point returns DRYV_ERR_CUDA.
#include <cuda_runtime.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <string>
#include <vector>

#include "recon_kernels.cuh"
#include "deblock_kernel.cuh"

namespace dryv {

__device__ __forceinline__ uint32_t ld_relaxed_gpu_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// lanes (field, macroblock) fetch mb_type / transform_size_8x8_flag / intra_chroma_pred_mode / qp of the macroblocks
// of a group; `field` picks the lane's array once per kernel
__device__ __forceinline__ const uint8_t* header_base(const KernelArgs& a, int field) {
  const uint8_t* p = a.qp;
  if (field == 0) p = a.mb_type;
  if (field == 1) p = a.t8x8;
  if (field == 2) p = a.chroma_mode;
  return p;
}

// Timeline trace (development builds only: -DDRYV_TRACE): the pixel warp records, for picture 0, the global
// timer at four points of every macroblock (slot ready, lines ready, prediction done, macroblock done).
#ifdef DRYV_TRACE
__device__ __forceinline__ unsigned int gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return (unsigned int)t;
}
#define TRACE_MARK(k)                                                                                  \
  do {                                                                                                 \
    if (a.trace && lane == 0 && G.frame == 0) a.trace[((size_t)row * W + x) * 4 + (k)] = gtimer_ns(); \
  } while (0)
#else
#define TRACE_MARK(k)
#endif

// Stage clocks (development builds only: -DDRYV_STAGE_CLOCKS): per-role cycle sums per stage.
#ifdef DRYV_STAGE_CLOCKS
#define CLK_DECL long long clk_t0 = clock64(), clk_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define CLK_MARK(i)                    \
  do {                                 \
    long long t_ = clock64();          \
    clk_acc[i] += t_ - clk_t0;         \
    clk_t0 = t_;                       \
  } while (0)
#define CLK_FLUSH(base)                                                                         \
  do {                                                                                          \
    if (lane == 0 && a.prof)                                                                    \
      for (int i_ = 0; i_ < 8; i_++) atomicAdd(a.prof + (base) + i_, (unsigned long long)clk_acc[i_]); \
  } while (0)
#else
#define CLK_DECL
#define CLK_MARK(i)
#define CLK_FLUSH(base)
#endif

// ------------------------------------------------------------------------------------------------
// mbarrier helpers (CTA-scope producer/consumer hand-off between the two warps of a row team)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// Bulk copy global -> shared (TMA engine, no tensor map: the bytes are contiguous): one lane arms the mbarrier with the
// byte count and issues the copy; the consumers wait on the mbarrier. 16-byte aligned addresses, size a multiple of 16.
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* bar) {
  const uint32_t d = smem_u32(smem_dst), b = smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
               "l"(gmem_src), "r"(bytes), "r"(b)
               : "memory");
}
// Ampere-style alternative for the level fetch (development knob DRYV_LEVEL_CPASYNC): every lane copies 16 bytes per
// instruction, completion through cp.async groups
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#ifndef DRYV_LEVEL_CPASYNC
#define DRYV_LEVEL_CPASYNC 0
#endif
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
  const uint32_t addr = smem_u32(b);
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(addr),
      "r"(parity), "r"(20000u)  // suspend-time hint (ns): sleep in hardware instead of spinning
      : "memory");
}

// Hand-off between the two warps of a team:
//   full[i]  (front -> pixel)  an mbarrier. The front warp normally runs ahead, so the pixel warp's
//            try_wait succeeds on its first issue.
//   empty[i] (pixel -> front)  a named barrier (ids 1..kSlots): the pixel warp announces with bar.arrive, the
//            front warp parks in bar.sync. The front warp's waits are long (a whole macroblock time); a warp
//            parked in bar.sync issues nothing, whereas an mbarrier try_wait / nanosleep loop was measured
//            re-issuing hundreds of times per macroblock in this kernel (sleepers are woken far earlier than
//            their time-out). Ids are compile-time constants: the SM has 64 barriers, a CTA that indexes them
//            with a register is charged all 16 and caps residency at 4 teams per SM.
#define DRYV_BAR_CASES(OP)                                         \
  switch (si) {                                                    \
    case 0: asm volatile(OP " 1, 64;" ::: "memory"); break;        \
    case 1: asm volatile(OP " 2, 64;" ::: "memory"); break;        \
    case 2: asm volatile(OP " 3, 64;" ::: "memory"); break;        \
    case 3: asm volatile(OP " 4, 64;" ::: "memory"); break;        \
    case 4: asm volatile(OP " 5, 64;" ::: "memory"); break;        \
    case 5: asm volatile(OP " 6, 64;" ::: "memory"); break;        \
    case 6: asm volatile(OP " 7, 64;" ::: "memory"); break;        \
    case 7: asm volatile(OP " 8, 64;" ::: "memory"); break;        \
    case 8: asm volatile(OP " 9, 64;" ::: "memory"); break;        \
    case 9: asm volatile(OP " 10, 64;" ::: "memory"); break;       \
    case 10: asm volatile(OP " 11, 64;" ::: "memory"); break;      \
    case 11: asm volatile(OP " 12, 64;" ::: "memory"); break;      \
    case 12: asm volatile(OP " 13, 64;" ::: "memory"); break;      \
    case 13: asm volatile(OP " 14, 64;" ::: "memory"); break;      \
    default: asm volatile(OP " 15, 64;" ::: "memory"); break;      \
  }
__device__ __forceinline__ void bar_sync_empty(unsigned si) { DRYV_BAR_CASES("bar.sync") }
__device__ __forceinline__ void bar_arrive_empty(unsigned si) { DRYV_BAR_CASES("bar.arrive") }
#undef DRYV_BAR_CASES
static_assert(kGroupSlots >= 2 && kGroupSlots * kTeamsPerCta <= 15, "named-barrier ids 1..15: one per team and group slot");

// Long waits (the front warp runs kSlots macroblocks ahead and then blocks on the pixel warp for a whole
// macroblock time): the hinted try_wait above is woken by every mbarrier event of the SM and re-issues ~100
// times per wait, so poll with a plain timed sleep instead — nothing on the critical path depends on how
// quickly a freed slot is noticed.
#ifndef DRYV_FRONT_SLEEP_NS
#define DRYV_FRONT_SLEEP_NS 500
#endif
__device__ __forceinline__ bool mbar_test(unsigned long long* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* b, uint32_t parity) {
  while (!mbar_test(b, parity)) __nanosleep(DRYV_FRONT_SLEEP_NS);
}

// Waits until the words of the lanes with `mine` set carry this launch's tag; returns the lane's payload.
// `first` is the value of a load issued earlier (so its latency overlapped with other work).
// On a watchdog trip `dead` is set and every later wait returns immediately (the kernel drains with
// garbage and the host reports DRYV_ERR_WATCHDOG).
#ifndef DRYV_LINE_SLEEP_NS
#define DRYV_LINE_SLEEP_NS 0u
#endif
#ifndef DRYV_LINE_LONG_NS
#define DRYV_LINE_LONG_NS 500u
#endif
#ifndef DRYV_POLL_PACE
#define DRYV_POLL_PACE 0
#endif
#ifndef DRYV_POLL_UNROLL
#define DRYV_POLL_UNROLL 1
#endif
// Measured (64 x 1080p): rolling the poll loop (DRYV_POLL_UNROLL 1) changes nothing; making wait_line_words a real
// function (__noinline__, -280 static instructions) costs 11 % (0.90 -> 1.00 ms): the call sits on the critical path.
#ifndef DRYV_WAIT_INLINE
#define DRYV_WAIT_INLINE __forceinline__
#endif
constexpr int kPollPace = DRYV_POLL_PACE;
constexpr int kPollUnroll = DRYV_POLL_UNROLL;
__device__ DRYV_WAIT_INLINE uint32_t wait_line_words_ns(const unsigned long long* p, unsigned long long first, bool mine,
                                                       uint32_t tag, unsigned ns, int* status, bool& dead, uint32_t& pace_addr);
__device__ DRYV_WAIT_INLINE uint32_t wait_line_words(const unsigned long long* p, unsigned long long first, bool mine,
                                                    uint32_t tag, bool long_wait, int* status, bool& dead,
                                                    uint32_t& pace_addr) {
  return wait_line_words_ns(p, first, mine, tag, long_wait ? DRYV_LINE_LONG_NS : DRYV_LINE_SLEEP_NS, status, dead, pace_addr);
}
//   ns: sleep between polls (0: spin)
__device__ DRYV_WAIT_INLINE uint32_t wait_line_words_ns(const unsigned long long* p, unsigned long long first, bool mine,
                                                       uint32_t tag, unsigned ns, int* status, bool& dead, uint32_t& pace_addr) {
  unsigned long long v = first;
  if (dead || __all_sync(0xffffffffu, !mine || (uint32_t)(v >> 32) == tag)) return (uint32_t)v;
  // Slow path. Only the lanes that own a word poll, each in its own loop (load, compare, branch: three instructions per
  // poll, no vote); the others wait at the __syncwarp below. Every instruction a waiting warp issues is taken from the
  // working warps of its scheduler, so the loop is kept this small and the watchdog counts in steps of eight polls.
  bool tripped = false;
  if (mine) {
    unsigned spins = 0;
#pragma unroll 1
    for (;;) {
#pragma unroll kPollUnroll
      for (int k = 0; k < 8; k++) {
        if (ns) __nanosleep(ns);
        if (kPollPace > 0) {
          // Optional pacing (development knob): a chain of dependent shared-memory loads on a word that holds its own
          // address. Measured: it does not help.
          uint32_t a = pace_addr;
#pragma unroll
          for (int i = 0; i < kPollPace; i++) asm volatile("ld.volatile.shared.u32 %0, [%0];" : "+r"(a) : : "memory");
          pace_addr = a;
        }
        v = ld_relaxed_gpu_u64(p);
        if ((uint32_t)(v >> 32) == tag) goto arrived;
      }
      spins += 8;
      if (spins > (1u << 22)) {  // watchdog: far beyond any legitimate wait
        atomicExch(status, STATUS_WATCHDOG);
        tripped = true;
        break;
      }
    }
  }
arrived:
  __syncwarp();
  if (__any_sync(0xffffffffu, tripped)) dead = true;
  return (uint32_t)v;
}

// ------------------------------------------------------------------------------------------------
// Prediction-mode derivation for Intra4x4 / Intra8x8 macroblocks, pred4x4.rs:363-427 / pred8x8.rs:698-764.
//
// Pixel-independent, so it runs as a pre-pass. A macroblock is a 4x4 grid of cells; a cell's mode needs the
// cell to its left and the cell above. An Intra8x8 macroblock stores each block's mode in the four cells it
// covers, which makes "A is Intra8x8 -> its 8x8 mode" and "A is Intra4x4 -> block 4*blk8+1" (and the B
// rules) plain cell look-ups; cells of Intra16x16 macroblocks hold 2 (DC), which is what the reference
// substitutes for a non-NxN neighbour.
//
// One CTA per picture, one lane per macroblock row, one warp per band of 32 rows. At its step s a lane
// resolves all sixteen cells of macroblock s - lane in registers (left column carried over from its previous
// macroblock), taking the bottom cell row of the macroblock above from the lane below it in index, which
// resolved it one step earlier (__shfl_up). Lane 0 reads it from the last row of the band above through
// shared memory, behind a progress counter. No block-wide barrier inside the walk.
// ------------------------------------------------------------------------------------------------
constexpr int kModeThreadsMax = 1024;  // 32 bands of 32 macroblock rows: pic_height_in_mbs <= 1024
__global__ void __launch_bounds__(kModeThreadsMax) resolve_modes_kernel(const KernelArgs a) {
  __shared__ uint16_t bandline[2][1024];  // bottom cell row of the last MB row of the even / odd bands, per MB column
  __shared__ int bandprog[32];            // columns finished by that row
  const int W = a.W, H = a.H;
  const size_t n_mb = (size_t)W * H;
  const size_t frame_mb0 = (size_t)blockIdx.x * n_mb;
  const int lane = threadIdx.x & 31, band = threadIdx.x >> 5;
  // Programmatic dependent launch: the wavefront kernel may start as soon as every CTA of this grid is resident;
  // it consumes the mode records as they appear (each one carries the launch tag).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#ifdef DRYV_TRACE
  if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[(size_t)a.W * a.H * 4] = gtimer_ns();
#endif
  if (threadIdx.x < 32) bandprog[threadIdx.x] = 0;
  // pull the picture's syntax (16 + 2 bytes per macroblock) into L2 up front: the walk below is a latency
  // chain with two macroblocks of register look-ahead, which covers an L2 hit but not a DRAM miss
  {
    const uint8_t* ps = a.pred_syntax + frame_mb0 * 16;
    for (size_t o = (size_t)threadIdx.x * 128; o < n_mb * 16; o += (size_t)blockDim.x * 128) prefetch_l2(ps + o);
    for (size_t o = (size_t)threadIdx.x * 128; o < n_mb; o += (size_t)blockDim.x * 128) {
      prefetch_l2(a.mb_type + frame_mb0 + o);
      prefetch_l2(a.t8x8 + frame_mb0 + o);
    }
  }
  __syncthreads();
  const int y = 32 * band + lane;  // macroblock row
  const bool row_ok = y < H;
  const size_t mb_row0 = frame_mb0 + (size_t)(row_ok ? y : 0) * W;
  const int rows_here = min(32, H - 32 * band);
  if (rows_here <= 0) return;
  const bool last_row = lane == rows_here - 1 && 32 * (band + 1) < H;
  volatile int* prog_above = &bandprog[band > 0 ? band - 1 : 0];
  const volatile uint16_t* line_above = bandline[(band + 1) & 1];
  // two macroblocks of look-ahead on the syntax loads (raw loaded values are only combined when consumed)
  uint4 syn_n1 = make_uint4(0, 0, 0, 0), syn_n2 = syn_n1;
  uint8_t mbt_n1 = 0, mbt_n2 = 0, t8_n1 = 0, t8_n2 = 0;
  auto load_mb = [&](int x, uint4& syn, uint8_t& mbt, uint8_t& t8) {
    if (row_ok && x >= 0 && x < W) {
      const size_t mb = mb_row0 + x;
      syn = __ldg(reinterpret_cast<const uint4*>(a.pred_syntax + mb * 16));
      mbt = __ldg(a.mb_type + mb);
      t8 = __ldg(a.t8x8 + mb);
    }
  };
  uint32_t leftcol = 0x2222u;  // modes of cells (3, 0..3) of this lane's previous macroblock, one nibble each
  uint32_t bottom = 0x2222u;   // modes of cells (0..3, 3) of the macroblock this lane resolved in the previous step
  const bool haveB0 = y > 0;
  const int steps = W + rows_here - 1;
  for (int s = -2; s < steps; s++) {
    const int x = s - lane;
    const uint4 syn_c = syn_n1;
    const int mbt = mbt_n1, t8 = t8_n1;
    syn_n1 = syn_n2;
    mbt_n1 = mbt_n2;
    t8_n1 = t8_n2;
    load_mb(x + 2, syn_n2, mbt_n2, t8_n2);
    if (s < 0) continue;
    uint32_t top4 = __shfl_up_sync(0xffffffffu, bottom, 1);
    if (lane == 0) {
      top4 = 0x2222u;
      if (band > 0 && x < W) {
        while (*prog_above <= x) __nanosleep(64);
        top4 = line_above[x];
      }
    }
    if (row_ok && x >= 0 && x < W) {
      const int cls = mbt == 0 ? (t8 ? 1 : 0) : 2;  // slice/macroblock.rs:682-716
      const bool haveA0 = x > 0;
      uint32_t c[4][4];  // resolved modes, [gy][gx]
      if (cls == 2) {
#pragma unroll
        for (int i = 0; i < 16; i++) c[i >> 2][i & 3] = 2;
      } else if (cls == 0) {
        const uint32_t sw[4] = {syn_c.x, syn_c.y, syn_c.z, syn_c.w};
#pragma unroll
        for (int gy = 0; gy < 4; gy++)
#pragma unroll
          for (int gx = 0; gx < 4; gx++) {
            const int blk = 8 * (gy >> 1) + 4 * (gx >> 1) + 2 * (gy & 1) + (gx & 1);  // pred4x4.rs:14-17
            const uint32_t syn = (sw[blk >> 2] >> (8 * (blk & 3))) & 0xffu;
            const uint32_t prev = (syn >> 3) & 1u, rem = syn & 7u;
            const uint32_t am = gx > 0 ? c[gy][gx - 1] : ((leftcol >> (4 * gy)) & 15u);
            const uint32_t bm = gy > 0 ? c[gy - 1][gx] : ((top4 >> (4 * gx)) & 15u);
            const bool haveA = gx > 0 || haveA0, haveB = gy > 0 || haveB0;
            const uint32_t pred = (haveA && haveB) ? min(am, bm) : 2u;  // pred4x4.rs:386-414
            c[gy][gx] = prev ? pred : (rem < pred ? rem : rem + 1);     // pred4x4.rs:416-426
          }
      } else {
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const int bx = b & 1, by = b >> 1;
          const uint32_t syn = (syn_c.x >> (8 * b)) & 0xffu;
          const uint32_t prev = (syn >> 3) & 1u, rem = syn & 7u;
          const uint32_t am = bx > 0 ? c[2 * by][1] : ((leftcol >> (8 * by)) & 15u);
          const uint32_t bm = by > 0 ? c[1][2 * bx] : ((top4 >> (8 * bx)) & 15u);
          const bool haveA = bx > 0 || haveA0, haveB = by > 0 || haveB0;
          const uint32_t pred = (haveA && haveB) ? min(am, bm) : 2u;  // pred8x8.rs:723-763
          const uint32_t m = prev ? pred : (rem < pred ? rem : rem + 1);
          c[2 * by][2 * bx] = c[2 * by][2 * bx + 1] = c[2 * by + 1][2 * bx] = c[2 * by + 1][2 * bx + 1] = m;
        }
      }
      leftcol = c[0][3] | (c[1][3] << 4) | (c[2][3] << 8) | (c[3][3] << 12);
      bottom = c[3][0] | (c[3][1] << 4) | (c[3][2] << 8) | (c[3][3] << 12);
      if (cls != 2) {
        // record layout: see kModeWords (recon_kernels.cuh)
        uint32_t lo = 0, hi = c[3][2] | (c[3][3] << 4);
#pragma unroll
        for (int gy = 0; gy < 4; gy++) lo |= (c[gy][0] | (c[gy][1] << 4)) << (8 * gy);
#pragma unroll
        for (int gy = 0; gy < 3; gy++) hi |= (c[gy][2] | (c[gy][3] << 4)) << (8 * (gy + 1));
        unsigned long long* rec = a.modes + (mb_row0 + x) * kModeWords;
        const unsigned long long tg = (unsigned long long)a.tag << 32;
        st_relaxed_gpu_u64(rec, tg | lo);
        st_relaxed_gpu_u64(rec + 1, tg | hi);
      }
#ifdef DRYV_TRACE
      if (a.trace && blockIdx.x == 0 && y == H - 1 && x == W - 1) a.trace[(size_t)a.W * a.H * 4 + 1] = gtimer_ns();
#endif
      if (last_row) {
        bandline[band & 1][x] = (uint16_t)bottom;
        __threadfence_block();
        *reinterpret_cast<volatile int*>(&bandprog[band]) = x + 1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Full reconstruction: persistent row teams over an x+2y macroblock wavefront.
//
// A row team (one CTA, two warps) walks one macroblock row of one picture left to right, a group of four macroblocks
// at a time:
//   front warp  - one bulk copy per group for the levels (a group ahead), the residual stage of the group
//                 (residual_stage.cuh) into a group slot, tap rows of the Intra4x4 blocks, then the chroma of each
//                 macroblock. Depends on nothing but its own macroblocks for luma, so it runs ahead of the pixels.
//   pixel warp  - Intra4x4/8x8/16x16 prediction, residual add + clip, 128-bit row stores.
// Rows hand data down through the line buffer: after a macroblock the team writes its bottom line
// (4 luma words, 2+2 chroma words), each as payload | launch tag in one 64-bit word. The row below
// fetches a line with one relaxed 64-bit load per lane, issued a whole macroblock before it is needed,
// and only checks the tags later: no fences, no flags, and nobody reads the picture back. Luma needs
// line x+1 of the row above (top-right neighbour: x+2y wavefront); chroma only needs line x.
// ------------------------------------------------------------------------------------------------
#if DRYV_CLUSTER > 1
// Cluster mode: entry `next` of the ring in the CTA below overwrites entry next - kRingEntries; wait until the row below
// has consumed it (`cons`: this CTA's counter, which the CTA below advances through distributed shared memory).
__device__ __noinline__ void ring_backpressure(volatile uint32_t* cons, unsigned next, unsigned& known, int lane, int* status,
                                               bool& dead) {
  bool trip = false;
  if (lane == 0 && !dead) {
    unsigned spins = 0;
    while (next - (known = *cons) >= (unsigned)kRingEntries) {
      __nanosleep(100);
      if (++spins > (1u << 22)) {
        atomicExch(status, STATUS_WATCHDOG);
        trip = true;
        break;
      }
    }
  }
  known = __shfl_sync(0xffffffffu, known, 0);
  if (__any_sync(0xffffffffu, trip)) dead = true;
}
#endif

// Start lag: how many macroblocks the row above must have finished before a row starts. The x+2y order needs
// two. Larger values were measured (3, 4, 6, 8): they lengthen the pipeline fill by (lag - 2) macroblock times
// per row and buy nothing, because per-macroblock cost varies by 3x between Intra16x16 and Intra4x4 and the
// slack is gone within a few macroblocks. Kept as a development knob.
#ifndef DRYV_CHROMA_LONG_WAIT
#define DRYV_CHROMA_LONG_WAIT 1
#endif
#ifndef DRYV_START_LAG
#define DRYV_START_LAG 2
#endif
constexpr int kStartLag = DRYV_START_LAG;
// Who turns an Intra4x4 macroblock's mode record into tap rows: the front warp (two macroblocks per pass) or the pixel warp
// (development knob). Measured, 64 x 1080p in flight / one at a time: pixel warp 0.718 / 0.827 (default 0.704 / 0.805) although
// the kernel shrinks by 112 static instructions — the pixel warp's time per macroblock is the hop of the dependency chain;
// together with DRYV_I4_UNROLL=2 and DRYV_LV_STAGES=1 (3096 instructions): 0.705 / 0.783, one picture 0.342 (0.353),
// Intra4x4-only batch 0.935 (0.970), 16 pictures in flight 0.270 (0.244): a wash, not adopted.
#ifndef DRYV_TAPROWS_BY_PIXEL
#define DRYV_TAPROWS_BY_PIXEL 0
#endif
static_assert(kCluster == 1 || (kStartLag == 2 && kTeamsPerCta == 1), "cluster mode: one team per CTA, default start lag");
// CTAs per SM = the register budget handed to ptxas; shared memory: 11 KB of tables per CTA + 14 KB per team.
#ifndef DRYV_CTAS_PER_SM
#define DRYV_CTAS_PER_SM 8
#endif
__global__ void __launch_bounds__(kWaveThreads, DRYV_CTAS_PER_SM) recon_wavefront_kernel(const KernelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WaveCtaSmem& cs = *reinterpret_cast<WaveCtaSmem*>(smem_raw);
  const unsigned team = threadIdx.x / kTeamThreads;
  TeamSmem& ts = cs.team[team];
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.tables);
    uint4* dst = reinterpret_cast<uint4*>(cs.tab);
    for (int i = threadIdx.x; i < (int)(kTeamTableBytes / 16); i += kWaveThreads) dst[i] = src[i];
    // tap-row bytes 16..31 of every slot stay zero: the look-ahead of the Intra4x4 loop reads up to byte 18
    const int tt = threadIdx.x % kTeamThreads;
    for (int i = tt; i < kGroupSlots * kGroupMbs * 4; i += kTeamThreads)
      reinterpret_cast<uint32_t*>(ts.grp[i / (kGroupMbs * 4)].mb[(i / 4) % kGroupMbs].rows + 16)[i % 4] = 0u;
    if (tt == 0) {
      for (int i = 0; i < kGroupSlots; i++) mbar_init(&ts.full[i], 1);
      for (int i = 0; i < kLvStages; i++) mbar_init(&ts.lvfull[i], 1);
      ts.pace = smem_u32(&ts.pace);  // a word that holds its own shared-memory address (see wait_line_words)
    }
#if DRYV_CLUSTER > 1
    for (int i = tt; i < kRingEntries * kLineWords; i += kTeamThreads) ts.ring[i] = 0ull;  // tag 0: no entry carries it
    if (tt < kMailSlots) ts.mail[tt] = 0ull;
    if (tt < 2) ts.cons[tt] = 0u;
    if (tt < kCluster) ts.ack[tt] = 0u;
#endif
  }
  __syncthreads();
#if DRYV_CLUSTER > 1
  // nobody writes into a neighbour's ring / mailbox before that neighbour has cleared it
  cluster_sync_all();
  const unsigned crank = cluster_ctarank();
  const bool ring_in = crank > 0;                         // the row above is walked by rank - 1 of this cluster
  const bool ring_out = crank + 1 < (unsigned)kCluster;   // the row below by rank + 1
  // this lane's word of ring entry 0, here and in the CTA that walks the row below; the counters the row above reads
  const unsigned long long* const ring_l = &ts.ring[threadIdx.x & 7];
  unsigned long long* const ringr_l = map_rank(&ts.ring[threadIdx.x & 7], ring_out ? crank + 1 : crank);
  volatile uint32_t* const cons_up = map_rank(&ts.cons[0], ring_in ? crank - 1 : crank);
#endif
  uint32_t pace_addr = smem_u32(&ts.pace);
  const int lane = threadIdx.x & 31;
#ifndef DRYV_ROLE_SWAP
#define DRYV_ROLE_SWAP 0
#endif
  // a CTA's first warp lands on schedulers 0 / 2 of the SM and its second on 1 / 3 (tools/micro/warp_slots.cu); with
  // DRYV_ROLE_SWAP every other CTA swaps the roles, so that front and pixel warps are spread over all four schedulers
  const bool is_front = (((threadIdx.x >> 5) & 1) ^ (DRYV_ROLE_SWAP ? (blockIdx.x & 1) : 0)) == 0;
  const DeviceTables& tab = *reinterpret_cast<const DeviceTables*>(cs.tab);  // everything but t4
  const unsigned bar0 = team * kGroupSlots;  // the team's named barriers
  const int W = a.W, H = a.H;
  const size_t n_mb = (size_t)W * H;
  const int strideY = W * 16;
  const uint32_t tag = a.tag;

  if (is_front) {
    // =========================================== front warp ===========================================
    const ResLane lc = make_res_lane(lane, tab);
    // header lanes: lane = field * 4 + macroblock of the group; field 0 mb_type, 1 transform_size_8x8_flag, 2 chroma mode, 3 qp
    const uint8_t* const hdr_base = header_base(a, lane >> 2);
    const uint32_t hdr_lim = (lane >> 2) == 0 ? 24u : ((lane >> 2) == 2 ? 3u : ((lane >> 2) == 3 ? 51u : 255u));
    const unsigned total_rows = (unsigned)a.n_frames * (unsigned)H;
    (void)total_rows;
    const int gpr = (W + kGroupMbs - 1) / kGroupMbs;
    bool unsupported = false;
    unsigned gn = 0;   // groups handed to the pixel warp so far
    unsigned lvw = 0;  // level fetches consumed so far: fetch k lands in stage k & 1, phase (k >> 1) & 1
    // chroma lane roles: lanes 4..5 / 6..7 fetch and publish the two Cb / Cr words of a bottom line,
    // lanes 8..9 shift the top-row slots [x-1].w1 | [x].w0..1 (byte 4 + 4k of tile row -1) when the walker
    // advances, lanes 16..23 / 24..31 store and carry one Cb / Cr pixel row each
    uint8_t* c_fresh = nullptr;
    const uint8_t* c_pub = nullptr;
    if (lane >= 4 && lane < 8) {
      const int pln = (lane - 4) >> 1, k = (lane - 4) & 1;
      c_fresh = &ts.chroma[pln * kChromaTileBytes + 4 + 4 * (1 + k)];
      c_pub = &ts.chroma[pln * kChromaTileBytes + chroma_at(4 * k, 7)];
    }
    uint8_t* const c_shift = (lane == 8 || lane == 9) ? &ts.chroma[(lane - 8) * kChromaTileBytes + 4] : nullptr;
    uint8_t* const c_tile = &ts.chroma[((lane >> 3) & 1) * kChromaTileBytes];
    const int c_first = chroma_at(0, lane & 7), c_last = chroma_at(7, lane & 7), c_left = chroma_at(-1, lane & 7);
    const int strideC = W * 8;
    const bool c_lane = lane >= 4 && lane < 8;
    // mode-record lanes: lanes 8..15 = (macroblock (lane >> 1) & 3 of the group, word lane & 1)
    const bool m_lane = lane >= 8 && lane < 16;
    const int m_mb = (lane >> 1) & 3;
    bool dead = false;
    CLK_DECL;
#if DRYV_CLUSTER > 1
    // A ticket is a band of kCluster consecutive rows of one picture (bands dealt band-major over pictures, so a band
    // only ever waits on a lower ticket); rank 0 draws it and posts it into the mailboxes of the other ranks.
    const unsigned bands = (unsigned)(H + kCluster - 1) / (unsigned)kCluster;
    const unsigned total_tickets = (unsigned)a.n_frames * bands;
    unsigned kt = 0;                       // tickets taken so far
    unsigned cin = 0, cout = 0, c_known = 0;  // chroma ring: entries consumed / published / known to be consumed below
    unsigned long long* const mail_r = map_rank(&ts.mail[0], (lane >= 1 && lane < kCluster) ? (unsigned)lane : crank);
    volatile uint32_t* const ack_r = map_rank(&ts.ack[crank], 0u);
#endif
    for (;;) {
      unsigned t = 0;
#if DRYV_CLUSTER > 1
      if (crank == 0) {
        if (kt >= (unsigned)kMailSlots) {  // slot kt & 3 still holds ticket kt - 4 until every rank has taken it
          bool trip = false;
          if (lane >= 1 && lane < kCluster) {
            unsigned spins = 0;
            while (ts.ack[lane] + (unsigned)kMailSlots - 1u < kt) {
              __nanosleep(200);
              if (++spins > (1u << 22)) {
                atomicExch(a.status, STATUS_WATCHDOG);
                trip = true;
                break;
              }
            }
          }
          if (__any_sync(0xffffffffu, trip)) dead = true;
        }
        if (lane == 0) t = atomicAdd(a.ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (lane >= 1 && lane < kCluster) st_relaxed_gpu_u64(mail_r + (kt & (kMailSlots - 1)), ((unsigned long long)(kt + 1u) << 32) | t);
      } else {
        if (lane == 0) {
          const unsigned long long* m = &ts.mail[kt & (kMailSlots - 1)];
          unsigned long long v = ld_relaxed_gpu_u64(m);
          unsigned spins = 0;
          while ((uint32_t)(v >> 32) != kt + 1u) {
            __nanosleep(200);
            v = ld_relaxed_gpu_u64(m);
            if (++spins > (1u << 22)) {
              atomicExch(a.status, STATUS_WATCHDOG);
              v = 0xffffffffull;  // no more work
              break;
            }
          }
          t = (uint32_t)v;
          st_relaxed_cluster_u32(ack_r, kt + 1u);
        }
        t = __shfl_sync(0xffffffffu, t, 0);
      }
      kt++;
      if (dead || t >= total_tickets) break;
      const int row = (int)(t / (unsigned)a.n_frames) * kCluster + (int)crank, frame = (int)(t % (unsigned)a.n_frames);
      if (row >= H) continue;  // the last band of a picture may be short
#else
      if (lane == 0) t = atomicAdd(a.ticket, 1u);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= total_rows) break;
      // tickets are dealt row-major over pictures so that a row only ever waits on a lower ticket
      const int row = (int)(t / (unsigned)a.n_frames), frame = (int)(t % (unsigned)a.n_frames);
#endif
      const size_t mb_row0 = (size_t)frame * n_mb + (size_t)row * W;
      const bool availB = row > 0, publish = row + 1 < H;
      const unsigned long long* c_above = a.line + (mb_row0 - W) * kLineWords + lane;  // chroma words of line x above
      unsigned long long* c_mine = a.line + mb_row0 * kLineWords + lane;
      uint8_t* c_st = a.out + (size_t)frame * n_mb * 384 + n_mb * 256 + (size_t)((lane >> 3) & 1) * n_mb * 64 +
                      (size_t)(8 * row + (lane & 7)) * strideC;
      unsigned long long lvc = 0;
#ifndef DRYV_CHROMA_LAG
#define DRYV_CHROMA_LAG 0
#endif
      if (DRYV_CHROMA_LAG > 0 && availB) {
        // Chroma start lag (development knob): the chroma walks of neighbouring rows run in lock-step (each macroblock waits
        // for the line the row above publishes a moment earlier); starting a row only when the row above is this many
        // macroblocks into its chroma walk gives the waits some slack to fall into.
        const unsigned long long* far = c_above + (size_t)min(DRYV_CHROMA_LAG, W - 1) * kLineWords;
        unsigned long long vf = 0;
        if (c_lane) vf = ld_relaxed_gpu_u64(far);
        wait_line_words(far, vf, c_lane, tag, true, a.status, dead, pace_addr);
      }
#if DRYV_CLUSTER > 1
      uint32_t c_tag = tag;
      if (ring_in) {
        c_above = ring_l + (size_t)(cin & (kRingEntries - 1)) * kLineWords;
        c_tag = cin + 1u;
      }
#else
      const uint32_t c_tag = tag;
#endif
      if (availB && c_lane) lvc = ld_relaxed_gpu_u64(c_above);

      const int16_t* const lv_row = a.coeff + mb_row0 * DRYV_COEFFS_PER_MB;
      auto fetch = [&](int g, unsigned k) {  // levels of group g of this row -> stage k & 1
        const int nn = min(kGroupMbs, W - g * kGroupMbs);
#if DRYV_LEVEL_CPASYNC
        const uint8_t* src = reinterpret_cast<const uint8_t*>(lv_row + (size_t)g * (kGroupMbs * DRYV_COEFFS_PER_MB));
        uint8_t* dst = reinterpret_cast<uint8_t*>(ts.lv[k % kLvStages]);
        for (int o = lane * 16; o < nn * (DRYV_COEFFS_PER_MB * 2); o += 512) cp_async_16(dst + o, src + o);
        cp_async_commit();
#else
        if (lane == 0)
          bulk_load(ts.lv[k % kLvStages], lv_row + (size_t)g * (kGroupMbs * DRYV_COEFFS_PER_MB), (uint32_t)nn * (DRYV_COEFFS_PER_MB * 2),
                    &ts.lvfull[k % kLvStages]);
#endif
      };
      auto load_hdr = [&](int g) -> uint32_t {
        const int x = g * kGroupMbs + (lane & 3);
        return (lane < 16 && x < W) ? (uint32_t)__ldg(hdr_base + mb_row0 + x) : 0u;
      };
      auto load_modes = [&](int g) -> unsigned long long {
        const int x = g * kGroupMbs + m_mb;
        return (m_lane && x < W) ? ld_relaxed_gpu_u64(a.modes + (mb_row0 + x) * kModeWords + (lane & 1)) : 0ull;
      };
      fetch(0, lvw);
      uint32_t hv = load_hdr(0);
      unsigned long long mv = load_modes(0);

      for (int g = 0; g < gpr; g++) {
        const int x0 = g * kGroupMbs, n = min(kGroupMbs, W - x0);
        const int stage = (int)(lvw % kLvStages);
#ifndef DRYV_FETCH_LATE
#define DRYV_FETCH_LATE 1
#endif
        if (!DRYV_FETCH_LATE && g + 1 < gpr) fetch(g + 1, lvw + 1);  // the other stage held group g - 1: consumed
        CLK_MARK(6);  // level fetch issued
        // ---- headers of the group ----
        uint32_t v = hv;
        const unsigned long long mvc = mv;
        if (g + 1 < gpr) {
          hv = load_hdr(g + 1);
          mv = load_modes(g + 1);
        }
        CLK_MARK(7);  // loads of the next group's headers / mode records issued
        if (v > hdr_lim) {  // I_PCM / inter / out-of-range syntax: flagged, never decoded
          unsupported = true;
          v = (lane >> 2) == 2 ? (v & 3u) : hdr_lim;
        }
        uint32_t m4, m8, mI4;
        {
          const int m = lane & 3;
          const uint32_t mbt = __shfl_sync(0xffffffffu, v, m), t8 = __shfl_sync(0xffffffffu, v, 4 + m),
                         cm = __shfl_sync(0xffffffffu, v, 8 + m), qp = __shfl_sync(0xffffffffu, v, 12 + m);
          const uint32_t cls = mbt == 0 ? (t8 ? 1u : 0u) : 2u;  // slice/macroblock.rs:682-716
          if (lane < kGroupMbs) ts.hdr[lane] = cls | (qp << 8) | (cm << 16) | (mbt << 24);
          m8 = __ballot_sync(0xffffffffu, lane < n && cls == 1u);
          mI4 = __ballot_sync(0xffffffffu, lane < n && cls == 0u);
          m4 = ((1u << n) - 1u) & ~m8;
        }
        __syncwarp();
        CLK_MARK(0);  // prefetch + headers
        // ---- group slot: wait until the pixel warp has released its previous use ----
        const unsigned gs = gn % kGroupSlots, use = gn / kGroupSlots;
        GroupSlot& G = ts.grp[gs];
        if (use > 0) bar_sync_empty(bar0 + gs);
        CLK_MARK(1);  // wait for a free slot
#if DRYV_LEVEL_CPASYNC
        if (g + 1 < gpr) cp_async_wait<1>();
        else cp_async_wait<0>();
        __syncwarp();
#else
        mbar_wait(&ts.lvfull[stage], (lvw / kLvStages) & 1u);
#endif
#ifndef DRYV_EXP_NO_RESID  // (experiment: how fast is each role with the other one's code out of the instruction cache)
        residual_group<false>(tab, a.tables, lc, lane, ts.hdr, m4, m8, ts.lv[stage], n, reinterpret_cast<int*>(&ts.cres[0][0]),
                       G.mb[0].res, (int)(sizeof(MbSlot) / sizeof(uint16_t)), &ts.cres[0][0], kResChromaMb, a.cb_off,
                       a.cr_off);
#endif
        lvw++;
        CLK_MARK(2);  // residual stage
        // ---- mode records of the Intra4x4 / Intra8x8 macroblocks (fetched a group ago): the pre-pass may still be
        // running, so check the tags and poll if a record has not got here yet (it walks a row about three times faster)
        if (m4 != ((1u << n) - 1u) || mI4) {
          const bool mine = m_lane && m_mb < n && (ts.hdr[m_mb] & 0xffu) != 2u;
          const uint32_t w = wait_line_words(a.modes + (mb_row0 + x0 + m_mb) * kModeWords + (lane & 1), mvc, mine, tag, true,
                                             a.status, dead, pace_addr);
          if (mine) (&G.mb[m_mb].modes_lo)[lane & 1] = w;
          __syncwarp();
          // Intra4x4: per-block tap rows (mode, top-right variant, legality and DC flavour in one byte), two macroblocks per pass
          if (!DRYV_TAPROWS_BY_PIXEL && mI4) {
            const uint32_t list = tab.setbits4[mI4];
            const int ni = __popc(mI4);
            for (int p = 0; 2 * p < ni; p++) {
              const uint32_t m = (list >> (4 * (2 * p + (lane >> 4)))) & 15u;
              if (m < (uint32_t)kGroupMbs) {
                const int x = x0 + (int)m;
                const int av = (x > 0 ? 1 : 0) | (availB ? 2 : 0) | ((availB && x + 1 < W) ? 4 : 0) | ((x > 0 && availB) ? 8 : 0);
                const uint2 mw = *reinterpret_cast<const uint2*>(&G.mb[m].modes_lo);
                G.mb[m].rows[lane & 15] = (uint8_t)i4_tap_row(tab, lane & 15, mw.x, mw.y, av);
              }
            }
          }
        }
        if (lane < n) {
          const uint32_t h = ts.hdr[lane];
          G.mb[lane].mbcls = (int)(h & 0xffu);
          G.mb[lane].mode16 = (int)(((h >> 24) - 1u) & 3u);
        }
        if (lane == 0) {
          G.frame = frame;
          G.row = row;
          G.x0 = x0;
          G.n = n;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ts.full[gs]);
        gn++;
        // The next group's levels: issuing the bulk copy costs the issuing warp ~1200 cycles (stage clocks), so it is issued
        // here, in front of the chroma walk, whose waits for the row above absorb it; the copy still has the whole chroma
        // phase of this group to land. (lvw already counts this group: the next fetch is number lvw, its stage the other one.)
        if (DRYV_FETCH_LATE && g + 1 < gpr) fetch(g + 1, lvw);
        CLK_MARK(4);  // hand-off

        // ---- chroma of the group's macroblocks: prediction needs line x of the row above (no top-right), so the chroma
        // walk of a row depends only on the chroma walk of the row above and stays off the luma critical path
        for (int m = 0; m < n; m++) {
          const int x = x0 + m;
          const int cm = (int)((ts.hdr[m] >> 16) & 0xffu);
          const bool availA = x > 0;
          if (availB) {
            // always the sleeping flavour: the front warp runs ahead of the pixel warp, so this wait is not on the
            // critical path, and a tight poll here would take issue slots from the pixel warps of the SM
            const uint32_t w = wait_line_words(c_above, lvc, c_lane, c_tag, DRYV_CHROMA_LONG_WAIT || x == 0, a.status, dead, pace_addr);
#if DRYV_CLUSTER > 1
            if (ring_in) {  // entry cin consumed: tell the row above, move to the next ring entry
              cin++;
              if (lane == 4) st_relaxed_cluster_u32(cons_up + 1, cin);
            }
#endif
            if (c_lane) {
              *reinterpret_cast<uint32_t*>(c_fresh) = w;
#if DRYV_CLUSTER > 1
              if (ring_in) {
                c_above = ring_l + (size_t)(cin & (kRingEntries - 1)) * kLineWords;
                c_tag = cin + 1u;
              } else
#endif
              c_above += kLineWords;
              if (x + 1 < W) lvc = ld_relaxed_gpu_u64(c_above);
            }
            __syncwarp();
          }
          CLK_MARK(3);  // wait for the chroma line of the row above
          predict_chroma(ts.chroma, ts.ccol, ts.cres[m], lane, cm, availA, availB, availA && availB);
#if DRYV_CLUSTER > 1
          if (publish && ring_out) {
            // entry cout of the ring of the row below; it overwrites entry cout - kRingEntries, which must have been consumed
            if (cout - c_known >= (unsigned)kRingEntries) ring_backpressure(&ts.cons[1], cout, c_known, lane, a.status, dead);
            if (c_lane)
              st_relaxed_gpu_u64(ringr_l + (size_t)(cout & (kRingEntries - 1)) * kLineWords,
                                 ((unsigned long long)(cout + 1u) << 32) | *reinterpret_cast<const uint32_t*>(c_pub));
            cout++;
          } else
#endif
          if (publish && c_lane)
            st_relaxed_gpu_u64(c_mine, ((unsigned long long)tag << 32) | *reinterpret_cast<const uint32_t*>(c_pub));
          c_mine += kLineWords;
          if (lane >= 16) {
            const uint2 pv = *reinterpret_cast<const uint2*>(&c_tile[c_first]);
            __stcs(reinterpret_cast<uint2*>(c_st), pv);
            c_st += 8;
          }
          {  // carry: right-most column -> left-neighbour column (tile column -1 and its contiguous copy), top-row shift
            const int cv = c_tile[c_last];
            uint32_t sv = 0;
            if (c_shift) sv = *reinterpret_cast<const uint32_t*>(c_shift + 8);
            __syncwarp();
            if (lane >= 16) {
              c_tile[c_left] = (uint8_t)cv;
              ts.ccol[lane - 16] = (uint8_t)cv;
            }
            if (c_shift) *reinterpret_cast<uint32_t*>(c_shift) = sv;
            __syncwarp();
          }
          CLK_MARK(5);  // chroma prediction + store + carry
        }
      }
    }
    CLK_FLUSH(0);
    // no more rows: tell the pixel warp
    {
      const unsigned gs = gn % kGroupSlots, use = gn / kGroupSlots;
      if (use > 0) bar_sync_empty(bar0 + gs);
      if (lane == 0) ts.grp[gs].row = -1;
      __syncwarp();
      if (lane == 0) mbar_arrive(&ts.full[gs]);
    }
    if (unsupported) atomicCAS(a.status, STATUS_OK, STATUS_UNSUPPORTED);
  } else {
    // =========================================== pixel warp ===========================================
    // Luma only. Top-row slots of the luma tile (tile row -1): 9 words [x-1].w3 | [x].w0..3 | [x+1].w0..3 at
    // byte 12 + 4k. Lanes 0..3 fetch the four luma words of line x+1 of the row above and publish this row's.
    const PixLane pl = make_pix_lane(lane);
    bool dead = false;
    uint8_t* const fresh_dst = lane < 4 ? &ts.luma[12 + 4 * (5 + lane)] : nullptr;       // where the fetched word goes
    const uint8_t* const pub_src = lane < 4 ? &ts.luma[luma_at(4 * lane, 15)] : nullptr; // the published word
    // top-row shift when the walker advances: lanes 0..4, slots k+4 -> k
    uint8_t* const shift_dst = lane < 5 ? &ts.luma[12 + 4 * lane] : nullptr;
    // lanes 0..15 store and carry one luma pixel row each
    const int my_first = luma_at(0, lane & 15), my_last = luma_at(15, lane & 15), my_left = luma_at(-1, lane & 15);
    unsigned gn = 0;
    const int W1 = W - 1;
    const unsigned long long* line_above = nullptr;  // + lane; line x+1 of the row above
    unsigned long long* line_mine = nullptr;
    uint8_t* st_ptr = nullptr;
    bool availB = false, publish = false;
    unsigned long long lv = 0;  // in-flight fetch of this lane's line word for the current macroblock
#if DRYV_CLUSTER > 1
    unsigned lj = 0;                  // luma ring: index of the entry line_above points at (entries below it are consumed)
    unsigned lout = 0, l_known = 0;   // entries published to the row below / known to be consumed there
    uint32_t l_tag = tag;
#else
    const uint32_t l_tag = tag;
#endif
    CLK_DECL;
    for (;;) {
      const unsigned gs = gn % kGroupSlots, use = gn / kGroupSlots;
      GroupSlot& G = ts.grp[gs];
      mbar_wait(&ts.full[gs], use & 1);
      CLK_MARK(0);  // wait for a filled slot
      const int row = G.row;
      if (row < 0) break;
      const int x0 = G.x0, n = G.n;
      for (int m = 0; m < n; m++) {
        const int x = x0 + m;
        MbSlot& slot = G.mb[m];
        if (x == 0) {
          // row start
          const int frame = G.frame;
          const size_t mb_row0 = (size_t)frame * n_mb + (size_t)row * W;
          availB = row > 0;
          publish = row + 1 < H;
          line_above = a.line + (mb_row0 - W) * kLineWords + lane;
#if DRYV_CLUSTER > 1
          if (ring_in) {
            line_above = ring_l + (size_t)(lj & (kRingEntries - 1)) * kLineWords;
            l_tag = lj + 1u;
          }
#endif
          line_mine = a.line + mb_row0 * kLineWords + lane;
          st_ptr = a.out + (size_t)frame * n_mb * 384 + (size_t)(16 * row + (lane & 15)) * strideY;
          if (availB) {
            // Start lag (see kStartLag), then luma line 0 of the row above becomes "line x" of macroblock 0
            if (kStartLag > 2) {
              const int ahead = min(kStartLag - 1, W1);
              const unsigned long long* far = line_above + (size_t)ahead * kLineWords;
              unsigned long long vf = 0;
              if (lane == 0) vf = ld_relaxed_gpu_u64(far);
              wait_line_words(far, vf, lane == 0, tag, true, a.status, dead, pace_addr);
            }
            unsigned long long v0 = 0;
            if (lane < 4) v0 = ld_relaxed_gpu_u64(line_above);
            const uint32_t w = wait_line_words(line_above, v0, lane < 4, l_tag, true, a.status, dead, pace_addr);
            if (lane < 4) *reinterpret_cast<uint32_t*>(fresh_dst) = w;
            __syncwarp();
            uint32_t sv = 0;
            if (lane < 5) sv = *reinterpret_cast<const uint32_t*>(shift_dst + 16);
            __syncwarp();
            if (lane < 5) *reinterpret_cast<uint32_t*>(shift_dst) = sv;
            // look one line ahead from here on
#if DRYV_CLUSTER > 1
            if (ring_in) {
              lj++;
              if (lane == 0) st_relaxed_cluster_u32(cons_up, lj);
              line_above = ring_l + (size_t)(lj & (kRingEntries - 1)) * kLineWords;
              l_tag = lj + 1u;
              if (lane < 4 && W1 > 0) lv = ld_relaxed_gpu_u64(line_above);
            } else
#endif
            if (lane < 4) {
              line_above += kLineWords;
              if (W1 > 0) lv = ld_relaxed_gpu_u64(line_above);
            }
          }
        }
        CLK_MARK(1);  // row start (incl. long wait for line 0)
        TRACE_MARK(0);
        const int mbcls = slot.mbcls, mode16 = slot.mode16;
        const bool availA = x > 0, availC = availB && x < W1, availD = availA && availB;
        // needs line x+1 of the row above (top-right neighbour), if it exists
        if (availC) {
          const uint32_t w = wait_line_words(line_above, lv, lane < 4, l_tag, false, a.status, dead, pace_addr);
          if (lane < 4) *reinterpret_cast<uint32_t*>(fresh_dst) = w;
#if DRYV_CLUSTER > 1
          if (ring_in) {  // entry lj consumed: tell the row above, move to the next ring entry (fetched after the prediction)
            lj++;
            if (lane == 0) st_relaxed_cluster_u32(cons_up, lj);
            line_above = ring_l + (size_t)(lj & (kRingEntries - 1)) * kLineWords;
            l_tag = lj + 1u;
          }
#endif
        }
        __syncwarp();
        CLK_MARK(2);  // wait for the luma line of the row above
        TRACE_MARK(1);

#ifdef DRYV_EXP_NO_PRED
        if (mbcls == 7) {
#else
        if (mbcls == 0) {
#endif
#if DRYV_TAPROWS_BY_PIXEL
          {  // the tap rows of the sixteen blocks: the pixel warp has the time (it waits for the front warp on average)
            const int av = (availA ? 1 : 0) | (availB ? 2 : 0) | (availC ? 4 : 0) | (availD ? 8 : 0);
            if (lane < 16) slot.rows[lane] = (uint8_t)i4_tap_row(tab, lane, slot.modes_lo, slot.modes_hi, av);
            __syncwarp();
          }
#endif
          predict_i4x4(tab, ts.luma, slot.res, pl, slot.rows);
          CLK_MARK(3);
#ifdef DRYV_EXP_NO_PRED
        } else if (mbcls == 8) {
#else
        } else if (mbcls == 1) {
#endif
          predict_i8x8(tab, ts.luma, ts.e8, slot.res, pl, lane, slot.modes_lo, slot.modes_hi, availA, availB, availC, availD);
          CLK_MARK(4);
        } else {
          predict_i16x16(ts.luma, ts.lcol, slot.res, lane, mode16, availA, availB);
          CLK_MARK(5);
        }
        // The row below is waiting for exactly this: publish the bottom line before anything else, and only then
        // fetch the line for the next macroblock (as late as possible: in a tightly coupled wavefront an earlier
        // load would only see that the row above has not got there yet).
#if DRYV_CLUSTER > 1
        if (publish && ring_out) {
          if (lout - l_known >= (unsigned)kRingEntries) ring_backpressure(&ts.cons[0], lout, l_known, lane, a.status, dead);
          if (lane < 4)
            st_relaxed_gpu_u64(ringr_l + (size_t)(lout & (kRingEntries - 1)) * kLineWords,
                               ((unsigned long long)(lout + 1u) << 32) | *reinterpret_cast<const uint32_t*>(pub_src));
          lout++;
        } else if (publish && lane < 4) {
          st_relaxed_gpu_u64(line_mine, ((unsigned long long)tag << 32) | *reinterpret_cast<const uint32_t*>(pub_src));
        }
        if (lane < 4 && availB) {
          if (!ring_in) line_above += kLineWords;  // (the ring pointer moved on when its entry was consumed)
          if (x + 2 <= W1) lv = ld_relaxed_gpu_u64(line_above);
        }
#else
        if (lane < 4) {
          if (publish)
            st_relaxed_gpu_u64(line_mine, ((unsigned long long)tag << 32) | *reinterpret_cast<const uint32_t*>(pub_src));
          if (availB) {
            line_above += kLineWords;
            if (x + 2 <= W1) lv = ld_relaxed_gpu_u64(line_above);
          }
        }
#endif
        line_mine += kLineWords;
        TRACE_MARK(2);

        // store the macroblock's 16 x 16 B luma rows
        if (lane < 16) {
          const uint4 pv = *reinterpret_cast<const uint4*>(&ts.luma[my_first]);
          __stcs(reinterpret_cast<uint4*>(st_ptr), pv);
          st_ptr += 16;
        }
        // carry: right-most column -> left-neighbour column (tile column -1 and its contiguous copy),
        // top-row slots shift by one macroblock
        const int cv = ts.luma[my_last];
        uint32_t sv = 0;
        if (shift_dst) sv = *reinterpret_cast<const uint32_t*>(shift_dst + 16);
        __syncwarp();
        if (lane < 16) {
          ts.luma[my_left] = (uint8_t)cv;
          ts.lcol[lane] = (uint8_t)cv;
        }
        if (shift_dst) *reinterpret_cast<uint32_t*>(shift_dst) = sv;
        __syncwarp();
        TRACE_MARK(3);
        CLK_MARK(7);  // stores + publish + carry
      }
      bar_arrive_empty(bar0 + gs);
      gn++;
    }
    CLK_FLUSH(8);
  }
#if DRYV_CLUSTER > 1
  // a CTA's shared memory (ring, mailbox, counters) must outlive every remote access to it
  __syncwarp();
  cluster_sync_all();
#endif
}

// ------------------------------------------------------------------------------------------------
// Residual only (BASELINE config "dequant + IDCT + residual add"): out = clip(pred_in + residual).
// No dependencies between macroblocks. One warp walks groups of four consecutive macroblocks (residual_stage.cuh):
//   * the group's 3072 B of levels arrive by ONE bulk copy (cp.async.bulk, mbarrier complete_tx) issued by one lane, one
//     group ahead of their use, into a two-stage ring;
//   * residual_group turns them into biased residual fields in shared memory (luma of two macroblocks / chroma of four
//     per pass, all 32 lanes busy);
//   * per macroblock one luma pass (8 samples per lane) and one chroma pass (4 samples per lane) add the prediction and
//     clip, two samples per VIADDMNMX, and store.
// ------------------------------------------------------------------------------------------------
struct ResidWarpSmem {
  alignas(16) int16_t lv[2][kGroupMbs * DRYV_COEFFS_PER_MB];  // level ring (bulk-copy destinations)
  alignas(16) uint16_t res_luma[kResLumaTile];                // residual fields of an Intra8x8 macroblock
  alignas(16) int scratch[kScratchWords];                     // 8x8 transposes
  alignas(16) uint32_t hdr[kGroupMbs];                        // class | qp << 8
  alignas(8) unsigned long long full[2];                      // mbarriers of the level ring
};
struct ResidCtaSmem {
  alignas(16) unsigned char tab[kResidTableBytes];  // the residual part of DeviceTables
  ResidWarpSmem warp[kWarpsPerCta];
};

// Position of a group: groups never straddle macroblock rows (the last group of a row is short when W % 4 != 0), so the
// macroblocks of a group are neighbours in the picture as well as in the level array.
struct GroupWalk {
  uint32_t gx, row, frame;  // group index inside the row, macroblock row, picture
  uint32_t W, H, gpr;
  __device__ __forceinline__ void init(uint32_t g, uint32_t W_, uint32_t H_) {
    W = W_;
    H = H_;
    gpr = (W + kGroupMbs - 1) / kGroupMbs;
    const uint32_t rowidx = g / gpr;
    gx = g - rowidx * gpr;
    frame = rowidx / H;
    row = rowidx - frame * H;
  }
  __device__ __forceinline__ void next() {
    if (++gx == gpr) {
      gx = 0;
      if (++row == H) {
        row = 0;
        frame++;
      }
    }
  }
  __device__ __forceinline__ uint32_t x0() const { return gx * kGroupMbs; }
  __device__ __forceinline__ int n() const { return (int)min((uint32_t)kGroupMbs, W - x0()); }
  __device__ __forceinline__ uint32_t mb0() const { return (frame * H + row) * W + x0(); }
};

#ifndef DRYV_RESID_CTAS
#define DRYV_RESID_CTAS 4
#endif
// FIELDS = true (split path, see recon_predict_kernel): nothing is added and no sample is stored; the macroblock's biased
// residual fields go to KernelArgs::resid instead, kResidMbFields per macroblock in the layout the predictors read from
// shared memory (luma tile of kResLumaTile fields, then the chroma tile of kResChromaMb fields), so that a row walker
// fetches them with one bulk copy.
template <bool FIELDS>
__device__ __forceinline__ void residual_kernel_body(const KernelArgs& a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ResidCtaSmem& cs = *reinterpret_cast<ResidCtaSmem*>(smem_raw);
  const int lane = threadIdx.x & 31;
  ResidWarpSmem& ws = cs.warp[threadIdx.x >> 5];
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.tables);
    uint4* dst = reinterpret_cast<uint4*>(cs.tab);
    for (int i = threadIdx.x; i < (int)(kResidTableBytes / 16); i += kThreadsPerCta) dst[i] = src[i];
    if (lane == 0) {
      mbar_init(&ws.full[0], 1);
      mbar_init(&ws.full[1], 1);
    }
  }
  __syncthreads();
  const DeviceTables& tab = *reinterpret_cast<const DeviceTables*>(cs.tab);  // only the residual part is there
  const ResLane lc = make_res_lane(lane, tab);
  const uint32_t W = (uint32_t)a.W, H = (uint32_t)a.H, n_mb = W * H;
  const uint32_t strideY = W * 16, strideC = W * 8;
  const uint32_t gpr = (W + kGroupMbs - 1) / kGroupMbs;
  const uint32_t groups = gpr * H * (uint32_t)a.n_frames;
  // every warp walks a contiguous range of groups
  const uint32_t warps = gridDim.x * kWarpsPerCta, wid = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  const uint32_t g_begin = (uint32_t)((unsigned long long)groups * wid / warps);
  const uint32_t g_end = (uint32_t)((unsigned long long)groups * (wid + 1) / warps);
  // header lanes: lane = field * 4 + macroblock of the group; field 0 mb_type, 1 transform_size_8x8_flag, 2 chroma mode, 3 qp
  const uint8_t* const hdr_base = header_base(a, lane >> 2);
  const uint32_t hdr_lim = (lane >> 2) == 0 ? 24u : ((lane >> 2) == 2 ? 3u : ((lane >> 2) == 3 ? 51u : 255u));
  bool unsupported = false;
  GroupWalk cur, nxt;
  cur.init(g_begin, W, H);
  nxt = cur;
  auto fetch = [&](const GroupWalk& p, int st) {
    bulk_load(ws.lv[st], a.coeff + (size_t)p.mb0() * DRYV_COEFFS_PER_MB, (uint32_t)p.n() * (DRYV_COEFFS_PER_MB * 2), &ws.full[st]);
  };
  auto load_hdr = [&](const GroupWalk& p) -> uint32_t {
    return (lane < 16 && (lane & 3) < p.n()) ? (uint32_t)__ldg(hdr_base + p.mb0() + (lane & 3)) : 0u;
  };
  uint32_t hv = 0;
  if (g_begin < g_end) {
    if (lane == 0) fetch(cur, 0);
    hv = load_hdr(cur);
  }
  for (uint32_t g = g_begin, it = 0; g < g_end; g++, it++) {
    const int st = it & 1;
    nxt.next();
    const bool more = g + 1 < g_end;
    if (more && lane == 0) fetch(nxt, st ^ 1);
    const int n = cur.n();
    uint32_t m4, m8;
    {  // headers of the group's macroblocks
      uint32_t v = hv;
      if (more) hv = load_hdr(nxt);  // next group's, one iteration ahead
      if (v > hdr_lim) {  // I_PCM / inter / out-of-range syntax: flagged, never decoded
        unsupported = true;
        v = (lane >> 2) == 2 ? (v & 3u) : hdr_lim;
      }
      const int m = lane & 3;
      const uint32_t mbt = __shfl_sync(0xffffffffu, v, m), t8 = __shfl_sync(0xffffffffu, v, 4 + m),
                     qp = __shfl_sync(0xffffffffu, v, 12 + m);
      const uint32_t cls = mbt == 0 ? (t8 ? 1u : 0u) : 2u;  // slice/macroblock.rs:682-716
      if (lane < kGroupMbs) ws.hdr[lane] = cls | (qp << 8);
      m8 = __ballot_sync(0xffffffffu, lane < n && cls == 1u);
      m4 = ((1u << n) - 1u) & ~m8;
      __syncwarp();
    }
    // Blocks with the 4x4 transform stay in registers from the levels to the stored samples: the lane that transforms a
    // block also adds its prediction and stores it (four 4-byte rows; the two macroblocks of a luma pass are neighbours,
    // so every 32-byte sector is still fully used). The prediction words of all passes are loaded here, before the
    // transforms, so that their latency is covered. Only Intra8x8 luma goes through a residual tile in shared memory.
    const size_t fo = (size_t)cur.frame * n_mb * 384;
    const uint8_t* const pin = a.pred_in + fo;
    uint8_t* const pout = a.out + fo;
    const size_t ybase = (size_t)(16u * cur.row) * strideY + 16u * cur.x0();               // luma of macroblock 0 of the group
    const size_t cbase = (size_t)n_mb * 256 + (size_t)(8u * cur.row) * strideC + 8u * cur.x0();  // Cb of macroblock 0
    const uint32_t list = tab.setbits4[m4];
    const int n4 = __popc(m4);
    size_t lo[2];        // this lane's block in luma pass 0 / 1: byte offset of its first row
    uint32_t pl[2][4];   // its prediction rows
    bool la[2];
#pragma unroll
    for (int p = 0; p < 2; p++) {
      const uint32_t m = (list >> (4 * (2 * p + (lane >> 4)))) & 15u;
      la[p] = 2 * p < n4 && m < (uint32_t)kGroupMbs;
      const int b4 = lane & 15;
      lo[p] = ybase + 16u * (la[p] ? m : 0u) + (size_t)((b4 >> 3) * 8 + ((b4 >> 1) & 1) * 4) * strideY + ((b4 >> 2) & 1) * 8 + (b4 & 1) * 4;
#pragma unroll
      for (int i = 0; i < 4; i++)
        pl[p][i] = (!FIELDS && la[p]) ? __ldg(reinterpret_cast<const uint32_t*>(pin + lo[p] + (size_t)i * strideY)) : 0u;
    }
    uint16_t* const fout = FIELDS ? a.resid + (size_t)cur.mb0() * kResidMbFields : nullptr;  // macroblock 0 of the group
    // chroma pass: lane = (macroblock lane >> 3, plane (lane >> 2) & 1, block lane & 3)
    const bool ca = (lane >> 3) < n;
    const size_t co = cbase + 8u * (ca ? (uint32_t)(lane >> 3) : 0u) + (size_t)((lane >> 2) & 1) * n_mb * 64 +
                      (size_t)(((lane >> 1) & 1) * 4) * strideC + (lane & 1) * 4;
    uint32_t pc[4];
#pragma unroll
    for (int i = 0; i < 4; i++) pc[i] = (!FIELDS && ca) ? __ldg(reinterpret_cast<const uint32_t*>(pin + co + (size_t)i * strideC)) : 0u;
    mbar_wait(&ws.full[st], (it >> 1) & 1);
    const int16_t* const lv = ws.lv[st];
    // ---- luma, 4x4 transform: two macroblocks per pass ----
#pragma unroll
    for (int p = 0; p < 2; p++) {
      if (2 * p < n4) {
        const uint32_t m = (list >> (4 * (2 * p + (lane >> 4)))) & 15u;
        const uint32_t mm = la[p] ? m : 0u;
        const uint32_t h = ws.hdr[mm];
        const int qp = (int)((h >> 8) & 0xffu);
        const bool i16 = (h & 0xffu) == 2u;
        const uint4* src = reinterpret_cast<const uint4*>(lv + mm * DRYV_COEFFS_PER_MB + (lane & 15) * 16);
        const uint4 c0 = src[0], c1 = src[1];
        int dcv = 0;
        if (__any_sync(0xffffffffu, la[p] && i16)) dcv = luma_dc16(tab, lc, lane, (int)(int16_t)(c0.x & 0xffffu), qp);
        uint32_t out[8];
        pass4x4_regs(tab, a.tables, c0, c1, qp, i16, dcv, la[p], out);
        if (FIELDS) {
          if (la[p]) {
            uint16_t* dst = fout + (size_t)mm * kResidMbFields + lc.res_off_luma;
#pragma unroll
            for (int i = 0; i < 4; i++) *reinterpret_cast<uint2*>(dst + i * kResLumaStride) = make_uint2(out[2 * i], out[2 * i + 1]);
          }
        } else if (la[p]) {
#pragma unroll
          for (int i = 0; i < 4; i++)
            __stcs(reinterpret_cast<uint32_t*>(pout + lo[p] + (size_t)i * strideY), add_clip4(out[2 * i], out[2 * i + 1], pl[p][i]));
        }
      }
    }
    // ---- chroma: the 8 blocks of each of the four macroblocks in one pass ----
    {
      const int mm = ca ? (lane >> 3) : 0;
      int q = (int)((ws.hdr[mm] >> 8) & 0xffu) + ((lane & 4) ? a.cr_off : a.cb_off);
      q = min(max(q, 0), 51);
      const int qpc = tab.qpc[q];  // transform.rs:194-216
      const uint4* src = reinterpret_cast<const uint4*>(lv + mm * DRYV_COEFFS_PER_MB + 256 + (lane & 7) * 16);
      const uint4 c0 = src[0], c1 = src[1];
      const int dcv = chroma_dc(tab, lane, (int)(int16_t)(c0.x & 0xffffu), qpc);
      uint32_t out[8];
      pass4x4_regs(tab, a.tables, c0, c1, qpc, true, dcv, ca, out);
      if (FIELDS) {
        if (ca) {
          uint16_t* dst = fout + (size_t)mm * kResidMbFields + kResLumaTile + lc.res_off_chroma;
#pragma unroll
          for (int i = 0; i < 4; i++) *reinterpret_cast<uint2*>(dst + i * 8) = make_uint2(out[2 * i], out[2 * i + 1]);
        }
      } else if (ca) {
#pragma unroll
        for (int i = 0; i < 4; i++)
          __stcs(reinterpret_cast<uint32_t*>(pout + co + (size_t)i * strideC), add_clip4(out[2 * i], out[2 * i + 1], pc[i]));
      }
    }
    // ---- luma, 8x8 transform: one macroblock per pass, through the residual tile; lane = (row r, half h): 8 samples ----
    for (uint32_t left = m8; left;) {
      const int m = __ffs(left) - 1;
      left &= left - 1;
      const size_t o = ybase + 16u * (uint32_t)m + (size_t)(lane >> 1) * strideY + 8 * (lane & 1);
      uint2 pv = make_uint2(0u, 0u);
      if (!FIELDS) pv = __ldg(reinterpret_cast<const uint2*>(pin + o));
      pass8x8(tab, lc, lane, lv + m * DRYV_COEFFS_PER_MB, ws.scratch, (int)((ws.hdr[m] >> 8) & 0xffu), ws.res_luma);
      if (FIELDS) {  // the tile as it lies in shared memory, 16 bytes per lane (kResLumaTile * 2 = 640 bytes)
        const uint4* src = reinterpret_cast<const uint4*>(ws.res_luma);
        uint4* dst = reinterpret_cast<uint4*>(fout + (size_t)m * kResidMbFields);
        dst[lane] = src[lane];
        if (lane < (int)(kResLumaTile * 2 / 16) - 32) dst[32 + lane] = src[32 + lane];
      } else {
        const uint2* rp = reinterpret_cast<const uint2*>(&ws.res_luma[(lane >> 1) * kResLumaStride + 8 * (lane & 1)]);
        const uint2 r0 = rp[0], r1 = rp[1];
        __stcs(reinterpret_cast<uint2*>(pout + o), make_uint2(add_clip4(r0.x, r0.y, pv.x), add_clip4(r1.x, r1.y, pv.y)));
      }
      __syncwarp();
    }
    __syncwarp();
    cur = nxt;
  }
  if (unsupported) atomicCAS(a.status, STATUS_OK, STATUS_UNSUPPORTED);
}
__global__ void __launch_bounds__(kThreadsPerCta, DRYV_RESID_CTAS) recon_residual_add_kernel(const KernelArgs a) {
  residual_kernel_body<false>(a);
}
__global__ void __launch_bounds__(kThreadsPerCta, DRYV_RESID_CTAS) recon_residual_fields_kernel(const KernelArgs a) {
  residual_kernel_body<true>(a);
}

#include "split_kernels.cuh"


// ------------------------------------------------------------------------------------------------
// Compact level stream -> dense level arrays (include/dryv_recon.h, dryv_mb_levels_compact): what the reference's
// residual_cabac does on the CPU when it scatters the significant levels into zero-filled block arrays
// (cabac/mod.rs:563-675). One warp per macroblock, lane b < 24 owns slot b (16 levels, 32 bytes): coded-slot mask ->
// index of the lane's significance mask, warp prefix sum of the population counts -> where the lane's levels start.
// `base` is the stream offset the device copy starts at (a chunk of a longer host stream), `stream_len` its length.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsPerCta) expand_levels_kernel(const uint32_t* __restrict__ offset,
                                                                      const uint8_t* __restrict__ stream, uint32_t base,
                                                                      uint32_t stream_len, size_t n_mbs,
                                                                      int16_t* __restrict__ coeff, int* status) {
  const int lane = threadIdx.x & 31;
  const size_t warps = (size_t)gridDim.x * kWarpsPerCta;
  bool bad = false;
  for (size_t mb = (size_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); mb < n_mbs; mb += warps) {
    const uint32_t o = __ldg(offset + mb) - base, e = __ldg(offset + mb + 1) - base;
    uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
    bool ok = e <= stream_len && o <= e && e - o >= 4u && e - o <= (uint32_t)DRYV_COMPACT_MAX_RECORD && !(o & 3u);
    if (ok) {
      const uint32_t len = e - o;
      const uint8_t* rec = stream + o;
      const uint32_t hdr = __ldg(reinterpret_cast<const uint32_t*>(rec));
      const uint32_t cm = hdr & 0xffffffu;
      const uint32_t mode = hdr >> 30;  // 0: int8 levels, 1: 4-bit codes + int16 escapes, 2: int16 levels
      const uint32_t ncoded = __popc(cm);
      const uint32_t lv0 = 4u + 2u * ncoded;  // where the levels start
      ok = !(hdr & 0x3f000000u) && mode <= 2u && lv0 <= len;
      uint32_t mask = 0;
      if (ok && lane < 24 && ((cm >> lane) & 1u))
        mask = __ldg(reinterpret_cast<const uint16_t*>(rec + 4 + 2 * __popc(cm & ((1u << lane) - 1u))));
      const uint32_t cnt = __popc(mask);
      uint32_t incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      const uint32_t first = incl - cnt;  // index of this lane's first level among the macroblock's levels
      uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (mode == 1u) {
        // 4-bit codes (bit 3 = sign, bits 0..2 = |level| 1..7, 0 = escape), two per byte, low nibble first, padded to
        // 2 bytes; the escaped levels follow as int16 in stream order
        const uint32_t nib_bytes = (((total + 1u) >> 1) + 1u) & ~1u;
        ok = ok && lv0 + nib_bytes <= len;
        uint32_t codes[16];
        uint32_t nesc = 0;
        if (ok) {
          const uint8_t* p = rec + lv0;
#pragma unroll
          for (int k = 0; k < 16; k++) {
            codes[k] = 0xffu;  // no level
            if ((mask >> k) & 1u) {
              const uint32_t j = first + __popc(mask & ((1u << k) - 1u));
              codes[k] = ((uint32_t)__ldg(p + (j >> 1)) >> (4u * (j & 1u))) & 15u;
              nesc += (codes[k] & 7u) == 0u;
            }
          }
        }
        uint32_t eincl = nesc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, eincl, d);
          if (lane >= d) eincl += t;
        }
        const uint32_t etotal = __shfl_sync(0xffffffffu, eincl, 31);
        ok = ok && lv0 + nib_bytes + 2u * etotal <= len;
        if (ok && mask) {
          const uint16_t* ep = reinterpret_cast<const uint16_t*>(rec + lv0 + nib_bytes) + (eincl - nesc);
#pragma unroll
          for (int k = 0; k < 16; k++) {
            int v = 0;
            if (codes[k] != 0xffu) {
              const uint32_t m3 = codes[k] & 7u;
              if (m3) v = (codes[k] & 8u) ? -(int)m3 : (int)m3;
              else v = (int)(int16_t)__ldg(ep++);
            }
            w[k >> 1] |= ((uint32_t)v & 0xffffu) << (16 * (k & 1));
          }
        }
      } else {
        const uint32_t wide = mode >> 1;
        ok = ok && lv0 + (total << wide) <= len;
        if (ok && mask) {
          const uint8_t* p = rec + lv0 + (first << wide);
#pragma unroll
          for (int k = 0; k < 16; k++) {
            // the k-th coefficient's level sits behind the levels of the set bits below k: independent loads
            int v = 0;
            if ((mask >> k) & 1u) {
              const uint32_t i = __popc(mask & ((1u << k) - 1u));
              v = wide ? (int)(int16_t)__ldg(reinterpret_cast<const uint16_t*>(p) + i) : (int)(int8_t)__ldg(p + i);
            }
            w[k >> 1] |= ((uint32_t)v & 0xffffu) << (16 * (k & 1));
          }
        }
      }
      if (ok) {
        c0 = make_uint4(w[0], w[1], w[2], w[3]);
        c1 = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
    if (!ok) bad = true;  // malformed record: the macroblock's levels are written as zeros and the launch is flagged
    if (lane < 24) {
      uint4* dst = reinterpret_cast<uint4*>(coeff + mb * DRYV_COEFFS_PER_MB) + lane * 2;
      dst[0] = c0;
      dst[1] = c1;
    }
  }
  if (bad) atomicCAS(status, STATUS_OK, STATUS_UNSUPPORTED);
}

// ------------------------------------------------------------------------------------------------
// Surface export: crop rectangle of the coded pictures -> packed I420 planes or NV12. A pure
// streaming pass: every thread moves V output bytes (V chosen on the host from the alignment of crop offset, width and
// plane offsets); rows of the coded picture that the rectangle leaves out are never read.
// ------------------------------------------------------------------------------------------------
template <int V> struct ExportVec;
template <> struct ExportVec<16> { typedef uint4 type; };
template <> struct ExportVec<8> { typedef uint2 type; };
template <> struct ExportVec<4> { typedef uint32_t type; };
template <> struct ExportVec<2> { typedef uint16_t type; };
template <> struct ExportVec<1> { typedef uint8_t type; };

// The whole surface in ONE launch (the three launches of the per-plane kernels left a fifth of the step to launch gaps):
// every thread moves V output bytes of one of the planes; a picture's threads are laid out luma | chroma, so a thread
// finds its plane by comparing its index inside the picture with the plane sizes. NV12: the chroma part interleaves
// V/2 Cb and V/2 Cr bytes with prmt.
struct ExportAllArgs {
  const uint8_t* src_y;   // first wanted luma sample of picture 0 (crop applied)
  const uint8_t* src_cb;  // ... Cb sample; Cr lies cr_delta bytes behind it
  uint8_t* dst;           // picture 0's surface
  size_t src_frame_stride, dst_frame_stride, cr_delta;
  uint32_t src_stride_y, src_stride_c;  // bytes between rows of the coded planes
  uint32_t w, h;                        // the rectangle, luma samples
  uint32_t vec_y, vec_c;                // vectors per picture: luma; one chroma plane (I420) / the interleaved plane (NV12)
  uint32_t nv12;
  unsigned long long total;             // n_frames * (vec_y + (nv12 ? 1 : 2) * vec_c)
};

template <int V>
__global__ void __launch_bounds__(256) export_all_kernel(const ExportAllArgs a) {
  typedef typename ExportVec<V>::type T;
  const unsigned long long idx = (unsigned long long)blockIdx.x * 256u + threadIdx.x;
  if (idx >= a.total) return;
  const uint32_t per_frame = a.vec_y + (a.nv12 ? 1u : 2u) * a.vec_c;
  const size_t f = (size_t)(idx / per_frame);
  uint32_t r = (uint32_t)(idx - (unsigned long long)f * per_frame);
  const uint8_t* sp = a.src_y + f * a.src_frame_stride;
  uint8_t* dp = a.dst + f * a.dst_frame_stride;
  if (r < a.vec_y) {
    const uint32_t per_row = a.w / V, row = r / per_row, c = r - row * per_row;
    __stcs(reinterpret_cast<T*>(dp + (size_t)row * a.w) + c, __ldcs(reinterpret_cast<const T*>(sp + (size_t)row * a.src_stride_y) + c));
    return;
  }
  r -= a.vec_y;
  dp += (size_t)a.w * a.h;
  sp = a.src_cb + f * a.src_frame_stride;
  if (!a.nv12) {
    if (r >= a.vec_c) {  // Cr
      r -= a.vec_c;
      sp += a.cr_delta;
      dp += (size_t)(a.w / 2) * (a.h / 2);
    }
    const uint32_t per_row = (a.w / 2) / V, row = r / per_row, c = r - row * per_row;
    __stcs(reinterpret_cast<T*>(dp + (size_t)row * (a.w / 2)) + c,
           __ldcs(reinterpret_cast<const T*>(sp + (size_t)row * a.src_stride_c) + c));
  } else if constexpr (V >= 2) {
    typedef typename ExportVec<V / 2>::type H;
    const uint32_t per_row = a.w / V, row = r / per_row, c = r - row * per_row;
    const size_t so = (size_t)row * a.src_stride_c;
    const H cb = __ldcs(reinterpret_cast<const H*>(sp + so) + c);
    const H cr = __ldcs(reinterpret_cast<const H*>(sp + a.cr_delta + so) + c);
    T out;
    if constexpr (V == 16) {
      out = make_uint4(__byte_perm(cb.x, cr.x, 0x5140), __byte_perm(cb.x, cr.x, 0x7362), __byte_perm(cb.y, cr.y, 0x5140),
                       __byte_perm(cb.y, cr.y, 0x7362));
    } else if constexpr (V == 8) {
      out = make_uint2(__byte_perm(cb, cr, 0x5140), __byte_perm(cb, cr, 0x7362));
    } else if constexpr (V == 4) {
      out = __byte_perm((uint32_t)cb, (uint32_t)cr, 0x5140);
    } else {
      out = (uint16_t)((uint32_t)cb | ((uint32_t)cr << 8));
    }
    __stcs(reinterpret_cast<T*>(dp + (size_t)row * a.w) + c, out);
  }
}

}  // namespace dryv

// ================================================================================================
// Host side: the C ABI
// ================================================================================================
using dryv::DeviceTables;
using dryv::KernelArgs;

#ifndef DRYV_SUBMIT_STAGES
#define DRYV_SUBMIT_STAGES 4
#endif
struct dryv_recon_ctx {
  int device = 0;
  int sm_count = 0;
  int wave_ctas_per_sm = 0, resid_ctas_per_sm = 0;
  int wave_grid_override = 0;  // DRYV_WAVE_GRID (development): row teams per launch instead of sm_count * teams per SM
  int wave_clusters = 0;  // cluster mode: clusters of the wavefront kernel that can be resident at once
  // s_compute[0] doubles as the default stream of the device-pointer entry points; dryv_recon_submit alternates
  // its chunks over both so that the (latency bound) wavefront kernels of neighbouring chunks overlap
  cudaStream_t s_compute[2] = {nullptr, nullptr}, s_h2d = nullptr, s_d2h = nullptr;
  static constexpr int kStages = DRYV_SUBMIT_STAGES;  // staging slots of the submit pipeline
  cudaEvent_t e_h2d[kStages] = {}, e_kernel[kStages] = {}, e_d2h[kStages] = {};
  cudaEvent_t e_sub_begin = nullptr, e_sub_end = nullptr;
  bool sub_timed = false;
  // submits may be queued back to back (the next batch's H2D runs under this batch's D2H): completion events of the
  // outstanding ones, oldest first, for dryv_recon_wait_oldest
  static constexpr int kPending = 4;
  cudaEvent_t e_done[kPending] = {};
  int pending_head = 0, pending_count = 0;
  bool sub_open = false;             // a submit has been queued since the last dryv_recon_wait
  bool slot_used[kStages] = {};      // the staging slot's events have been recorded at least once
  // CUDA-event pairs around the most recent wavefront-kernel launches (bench: roofline of the dominant kernel)
  static constexpr int kTimedLaunches = 64;
  cudaEvent_t e_wave[kTimedLaunches][2] = {};
  uint64_t wave_launches = 0;
  bool use_pdl = true;  // development switch: DRYV_NO_PDL=1 serialises the pre-pass and the wavefront kernel
  // DRYV_SPLIT=1: residual-fields kernel + one-warp-per-row predict kernel (split_kernels.cuh) instead of the row teams
  bool use_split = false;
  int split_ctas_per_sm = 0;
  unsigned int* d_ticket_c = nullptr;  // [set] chroma row tickets of the split path
  // tables
  DeviceTables* d_tables = nullptr;
  DeviceTables* h_tables = nullptr;  // pinned
  dryv_pic_params tables_pp;
  bool tables_valid = false;
  // Wavefront control blocks, used round robin: a launch may overlap the launches before it (on other streams), which
  // is where back-to-back batches gain — the start-up stagger of one wavefront fills the tail of the previous one
  // (64 x 1080p: 0.89 -> 0.74 ms per batch with two in flight, 8-picture batches 0.38 -> 0.13 ms with four). A block is
  // reused only after the launch that used it last has finished (its `done` event).
  static constexpr int kSets = 4;
  struct Control {
    unsigned long long* d_line = nullptr;  // bottom-line hand-off buffer, kLineWords words per macroblock
    unsigned long long* d_modes = nullptr; // resolved prediction modes, kModeWords tagged words per macroblock
    size_t line_cap = 0;                   // in macroblocks
    uint16_t* d_resid = nullptr;           // split path: kResidMbFields residual fields per macroblock
    size_t resid_cap = 0;
    cudaEvent_t done = nullptr;            // recorded behind the last launch that used this block
    bool used = false;
  } ctl[kSets];
  unsigned next_set = 0;
  uint32_t tag = 0;                      // launch tag, incremented per wavefront launch (0 = never written)
  unsigned int* d_ticket = nullptr;  // [2 * set] ticket of control block `set`, [1] status
  unsigned long long* d_prof = nullptr;  // stage clocks (development builds)
  unsigned int* d_trace = nullptr;       // timeline trace (development builds), 4 words per macroblock of picture 0
  size_t trace_mbs = 0;
  int* h_status = nullptr;           // pinned
  // staging for dryv_recon_submit / dryv_recon_submit_compact
  uint8_t* d_in[kStages] = {};
  uint8_t* d_out[kStages] = {};
  size_t in_cap = 0, out_cap = 0;
  // output surface of the submit calls (dryv_recon_set_surface): exported on the GPU into d_exp, copied out from there
  dryv_surface surface = {};
  bool surface_set = false;
  uint8_t* d_exp[kStages] = {};
  size_t exp_cap = 0;
  // deblocking post-pass: row ticket + the tagged words rows hand down; one launch of it at a time (guarded by db_done)
  unsigned int* d_db_ticket = nullptr;
  unsigned long long* d_db_line = nullptr;
  size_t db_line_cap = 0;  // macroblocks
  uint32_t db_tag = 0;
  // dryv_recon_set_deblock: the host submit paths run the post-pass behind the reconstruction of every slot
  bool deblock_on = false;
  int db_alpha_div2 = 0, db_beta_div2 = 0;
  cudaEvent_t db_done = nullptr;
  bool db_used = false;
  // launches on caller streams: one completion event per distinct stream, re-recorded by every launch on it, so that
  // dryv_recon_wait and a table re-upload cover all of them (an event outlives its stream)
  struct UserStream {
    cudaStream_t stream;
    cudaEvent_t done;
  };
  std::vector<UserStream> user_streams;
  uint64_t launches = 0;
  // development aid (DRYV_SUBMIT_TRACE=1): per-chunk stage completion events of the last submit, printed by wait
  std::vector<cudaEvent_t> trace_ev;
  std::string err;
};

namespace {

int fail(dryv_recon_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define CU(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      return fail(ctx, DRYV_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
  } while (0)

bool pp_ok(const dryv_pic_params* pp) {
  return pp && pp->pic_width_in_mbs >= 1 && pp->pic_width_in_mbs <= 1024 && pp->pic_height_in_mbs >= 1 &&
         pp->pic_height_in_mbs <= 1024 && pp->flags == 0 && pp->chroma_qp_index_offset >= -12 &&
         pp->chroma_qp_index_offset <= 12 && pp->second_chroma_qp_index_offset >= -12 &&
         pp->second_chroma_qp_index_offset <= 12;
}

// a launch went to the caller's stream `st`: remember to wait for it
int note_user_stream(dryv_recon_ctx* ctx, cudaStream_t st) {
  for (auto& u : ctx->user_streams)
    if (u.stream == st) {
      CU(cudaEventRecord(u.done, st));
      return DRYV_OK;
    }
  if (ctx->user_streams.size() >= 64) {  // a caller that keeps creating streams: retire the oldest entry
    CU(cudaEventSynchronize(ctx->user_streams.front().done));
    cudaEventDestroy(ctx->user_streams.front().done);
    ctx->user_streams.erase(ctx->user_streams.begin());
  }
  dryv_recon_ctx::UserStream u{st, nullptr};
  CU(cudaEventCreateWithFlags(&u.done, cudaEventDisableTiming));
  CU(cudaEventRecord(u.done, st));
  ctx->user_streams.push_back(u);
  return DRYV_OK;
}
int sync_user_streams(dryv_recon_ctx* ctx) {
  for (auto& u : ctx->user_streams) CU(cudaEventSynchronize(u.done));
  return DRYV_OK;
}

int ensure_tables(dryv_recon_ctx* ctx, const dryv_pic_params* pp, cudaStream_t s) {
  if (ctx->tables_valid && memcmp(&ctx->tables_pp, pp, sizeof *pp) == 0) return DRYV_OK;
  // the pinned host copy may still be in flight from a previous upload on another stream
  CU(cudaStreamSynchronize(ctx->s_compute[0]));
  CU(cudaStreamSynchronize(ctx->s_compute[1]));
  {
    const int rc = sync_user_streams(ctx);  // kernels on caller streams may still read the tables
    if (rc != DRYV_OK) return rc;
  }
  dryv::build_device_tables(*pp, ctx->h_tables);
  CU(cudaMemcpyAsync(ctx->d_tables, ctx->h_tables, sizeof(DeviceTables), cudaMemcpyHostToDevice, s));
  CU(cudaStreamSynchronize(s));
  ctx->tables_pp = *pp;
  ctx->tables_valid = true;
  return DRYV_OK;
}

int ensure_control(dryv_recon_ctx* ctx, int set, size_t mbs) {
  dryv_recon_ctx::Control& c = ctx->ctl[set];
  if (mbs > c.line_cap) {
    CU(cudaDeviceSynchronize());
    if (c.d_line) cudaFree(c.d_line);
    if (c.d_modes) cudaFree(c.d_modes);
    c.d_line = nullptr;
    c.d_modes = nullptr;
    c.line_cap = 0;
    const size_t bytes = mbs * dryv::kLineWords * sizeof(unsigned long long);
    CU(cudaMalloc(&c.d_line, bytes));
    CU(cudaMemset(c.d_line, 0, bytes));  // tag 0 is never used by a launch
    CU(cudaMalloc(&c.d_modes, mbs * dryv::kModeWords * sizeof(unsigned long long)));
    CU(cudaMemset(c.d_modes, 0, mbs * dryv::kModeWords * sizeof(unsigned long long)));  // tag 0 is never used
    c.line_cap = mbs;
  }
  return DRYV_OK;
}

KernelArgs make_args(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const dryv_mb_soa* soa, uint32_t n_frames,
                     uint8_t* out, int set = 0) {
  KernelArgs a;
  memset(&a, 0, sizeof a);
  a.mb_type = soa->mb_type;
  a.t8x8 = soa->transform_size_8x8_flag;
  a.chroma_mode = soa->intra_chroma_pred_mode;
  a.qp = soa->qp;
  a.pred_syntax = soa->pred_syntax;
  a.coeff = soa->coeff;
  a.out = out;
  a.tables = ctx->d_tables;
  a.line = ctx->ctl[set].d_line;
  a.modes = ctx->ctl[set].d_modes;
  a.tag = ctx->tag;
  a.ticket = ctx->d_ticket + 2 * set;
  a.status = reinterpret_cast<int*>(ctx->d_ticket + 1);
  a.ticket_c = ctx->d_ticket_c + set;
  a.resid = ctx->ctl[set].d_resid;
#ifdef DRYV_STAGE_CLOCKS
  a.prof = ctx->d_prof;
#endif
#ifdef DRYV_TRACE
  a.trace = ctx->d_trace;
#endif
  a.W = pp->pic_width_in_mbs;
  a.H = pp->pic_height_in_mbs;
  a.n_frames = (int)n_frames;
  a.cb_off = pp->chroma_qp_index_offset;
  a.cr_off = pp->second_chroma_qp_index_offset;
  return a;
}

bool soa_ok(const dryv_mb_soa* s) {
  return s && s->mb_type && s->transform_size_8x8_flag && s->intra_chroma_pred_mode && s->qp && s->pred_syntax &&
         s->coeff && (reinterpret_cast<uintptr_t>(s->coeff) % 16 == 0);
}

// enqueue the wavefront kernel for device-resident buffers on stream s
int launch_wavefront(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const dryv_mb_soa* d_soa, uint32_t n_frames,
                     uint8_t* d_out, cudaStream_t s) {
  const size_t rows = (size_t)n_frames * pp->pic_height_in_mbs;
  const size_t mbs = rows * pp->pic_width_in_mbs;
  const int set = (int)(ctx->next_set++ % dryv_recon_ctx::kSets);
  for (int i = 0; i < dryv_recon_ctx::kSets; i++) {  // all blocks grow together: no allocation once a batch size has run
    int rc = ensure_control(ctx, i, mbs);
    if (rc != DRYV_OK) return rc;
  }
  if (ctx->ctl[set].used) CU(cudaStreamWaitEvent(s, ctx->ctl[set].done, 0));  // the block's previous launch has finished
  if (++ctx->tag == 0) ctx->tag = 1;  // every launch validates line words with its own tag: no per-launch clearing
  CU(cudaMemsetAsync(ctx->d_ticket + 2 * set, 0, sizeof(unsigned int), s));  // ticket only; status stays sticky until wait
  if (ctx->use_split) {
    dryv_recon_ctx::Control& c = ctx->ctl[set];
    if (mbs > c.resid_cap) {
      CU(cudaDeviceSynchronize());
      if (c.d_resid) cudaFree(c.d_resid);
      c.d_resid = nullptr;
      c.resid_cap = 0;
      CU(cudaMalloc(&c.d_resid, mbs * dryv::kResidMbFields * sizeof(uint16_t)));
      c.resid_cap = mbs;
    }
    CU(cudaMemsetAsync(ctx->d_ticket_c + set, 0, sizeof(unsigned int), s));
  }
  KernelArgs a = make_args(ctx, pp, d_soa, n_frames, d_out, set);
  cudaEvent_t* ev = ctx->e_wave[ctx->wave_launches % dryv_recon_ctx::kTimedLaunches];
  CU(cudaEventRecord(ev[0], s));
  // pre-pass: prediction-mode derivation, one CTA per picture
  const int mode_threads = (((int)pp->pic_height_in_mbs + 31) / 32) * 32;  // one warp per band of 32 MB rows
  dryv::resolve_modes_kernel<<<n_frames, mode_threads, 0, s>>>(a);
  CU(cudaGetLastError());
  if (ctx->use_split) {
    // residual fields of every macroblock (no dependency on the pre-pass: a programmatic dependent of it, so the two overlap),
    // then the row walkers, which start once both have finished
    {
      const size_t gpr = (pp->pic_width_in_mbs + dryv::kGroupMbs - 1) / dryv::kGroupMbs;
      const size_t groups = gpr * rows;
      size_t want = (groups + dryv::kWarpsPerCta - 1) / dryv::kWarpsPerCta;
      size_t cap = (size_t)ctx->sm_count * ctx->resid_ctas_per_sm;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(want < cap ? want : cap));
      cfg.blockDim = dim3(dryv::kThreadsPerCta);
      cfg.dynamicSmemBytes = sizeof(dryv::ResidCtaSmem);
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = ctx->use_pdl ? 1 : 0;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      CU(cudaLaunchKernelEx(&cfg, dryv::recon_residual_fields_kernel, a));
    }
    {
      // one luma walker per macroblock row at most
      size_t want = (rows + dryv::kSplitLumaWarps - 1) / dryv::kSplitLumaWarps;
      size_t cap = (size_t)ctx->sm_count * ctx->split_ctas_per_sm;
      const unsigned grid = (unsigned)(want < cap ? want : cap);
      dryv::recon_predict_kernel<<<grid, dryv::kSplitThreads, sizeof(dryv::SplitCtaSmem), s>>>(a);
      CU(cudaGetLastError());
    }
    CU(cudaEventRecord(ev[1], s));
    CU(cudaEventRecord(ctx->ctl[set].done, s));
    ctx->ctl[set].used = true;
    ctx->wave_launches++;
    ctx->launches += 3;
    return DRYV_OK;
  }
  size_t want = (rows + dryv::kTeamsPerCta - 1) / dryv::kTeamsPerCta;  // one row team per macroblock row at most
  size_t cap = (size_t)ctx->sm_count * ctx->wave_ctas_per_sm;
#if DRYV_CLUSTER > 1
  // whole clusters only: one per band of kCluster rows at most, and no more than can be resident together
  want = (size_t)n_frames * ((pp->pic_height_in_mbs + dryv::kCluster - 1) / dryv::kCluster) * dryv::kCluster;
  cap = (size_t)ctx->wave_clusters * dryv::kCluster;
#endif
  int grid = (int)(want < cap ? want : cap);
  if (ctx->wave_grid_override > 0 && (size_t)ctx->wave_grid_override < cap) grid = ctx->wave_grid_override;  // development knob
  // The wavefront kernel is a programmatic dependent of the pre-pass: it starts once every pre-pass CTA is
  // resident and consumes mode records as they appear (tagged words, no grid-wide wait).
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(dryv::kWaveThreads);
  cfg.dynamicSmemBytes = sizeof(dryv::WaveCtaSmem);
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = ctx->use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#if DRYV_CLUSTER > 1
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = dryv::kCluster;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.numAttrs = 2;
#endif
  CU(cudaLaunchKernelEx(&cfg, dryv::recon_wavefront_kernel, a));
  CU(cudaEventRecord(ev[1], s));
  CU(cudaEventRecord(ctx->ctl[set].done, s));
  ctx->ctl[set].used = true;
  ctx->wave_launches++;
  ctx->launches += 2;
  return DRYV_OK;
}


// enqueue the compact-stream expansion for `n_mbs` macroblocks: d_stream points at stream byte `base`
int launch_expand(dryv_recon_ctx* ctx, const uint32_t* d_offset, const uint8_t* d_stream, uint32_t base, uint32_t len,
                  size_t n_mbs, int16_t* d_coeff, cudaStream_t s) {
  size_t want = (n_mbs + dryv::kWarpsPerCta - 1) / dryv::kWarpsPerCta;
  size_t cap = (size_t)ctx->sm_count * 16;
  int grid = (int)(want < cap ? want : cap);
  dryv::expand_levels_kernel<<<grid, dryv::kThreadsPerCta, 0, s>>>(d_offset, d_stream, base, len, n_mbs, d_coeff,
                                                                 reinterpret_cast<int*>(ctx->d_ticket + 1));
  CU(cudaGetLastError());
  ctx->launches++;
  return DRYV_OK;
}


bool surface_ok(const dryv_surface* s) {
  return s && (s->format == DRYV_SURFACE_I420 || s->format == DRYV_SURFACE_NV12) && s->width > 0 && s->height > 0 &&
         !((s->width | s->height | s->crop_left | s->crop_top) & 1u) && s->width <= (1u << 15) && s->height <= (1u << 15) &&
         s->crop_left <= (1u << 15) && s->crop_top <= (1u << 15);
}
size_t surface_bytes(const dryv_surface* s) { return (size_t)s->width * s->height * 3 / 2; }

// largest power of two <= cap that divides every one of the byte quantities
int common_vector(std::initializer_list<unsigned long long> q, int cap, int floor_v) {
  int v = cap;
  for (; v > floor_v; v >>= 1) {
    bool ok = true;
    for (unsigned long long x : q) ok = ok && (x % (unsigned)v) == 0;
    if (ok) break;
  }
  return v;
}

// enqueue the surface export of n_frames coded pictures (d_yuv, `pp` geometry) into d_out on stream st: one launch
int launch_export(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const uint8_t* d_yuv, uint32_t n_frames,
                  const dryv_surface* s, uint8_t* d_out, cudaStream_t st) {
  const size_t W = 16u * (size_t)pp->pic_width_in_mbs, H = 16u * (size_t)pp->pic_height_in_mbs;
  if ((size_t)s->crop_left + s->width > W || (size_t)s->crop_top + s->height > H)
    return fail(ctx, DRYV_ERR_ARG, "surface rectangle leaves the coded picture");
  const size_t src_frame = W * H * 3 / 2, dst_frame = surface_bytes(s);
  const size_t w = s->width, h = s->height, cl = s->crop_left, ct = s->crop_top;
  const unsigned long long pa = (unsigned long long)reinterpret_cast<uintptr_t>(d_yuv),
                           pb = (unsigned long long)reinterpret_cast<uintptr_t>(d_out);
  const bool nv12 = s->format == DRYV_SURFACE_NV12;
  // one vector width for every plane: it has to divide the luma quantities and, per chroma plane, half of them
  // (NV12: V output bytes = V/2 source bytes of each plane). W / 2: the coded chroma row stride is 8 * pic_width_in_mbs,
  // so with an odd macroblock count rows alternate between 16- and 8-byte alignment.
  const int v = nv12 ? common_vector({pa, pb, cl, w, w * h, dst_frame}, 16, 2)
                     : common_vector({pa, pb, cl / 2, w / 2, W / 2, w * h, w * h / 4, dst_frame}, 16, 1);
  dryv::ExportAllArgs a;
  memset(&a, 0, sizeof a);
  a.src_y = d_yuv + ct * W + cl;
  a.src_cb = d_yuv + W * H + (ct / 2) * (W / 2) + cl / 2;
  a.cr_delta = (W / 2) * (H / 2);
  a.dst = d_out;
  a.src_frame_stride = src_frame;
  a.dst_frame_stride = dst_frame;
  a.src_stride_y = (uint32_t)W;
  a.src_stride_c = (uint32_t)(W / 2);
  a.w = (uint32_t)w;
  a.h = (uint32_t)h;
  a.nv12 = nv12 ? 1u : 0u;
  a.vec_y = (uint32_t)(h * (w / v));
  a.vec_c = (uint32_t)(nv12 ? (h / 2) * (w / v) : (h / 2) * (w / 2 / v));
  a.total = (unsigned long long)n_frames * (a.vec_y + (nv12 ? 1u : 2u) * a.vec_c);
  if (a.total == 0) return DRYV_OK;
  const unsigned blocks = (unsigned)((a.total + 255) / 256);
  switch (v) {
    case 16: dryv::export_all_kernel<16><<<blocks, 256, 0, st>>>(a); break;
    case 8: dryv::export_all_kernel<8><<<blocks, 256, 0, st>>>(a); break;
    case 4: dryv::export_all_kernel<4><<<blocks, 256, 0, st>>>(a); break;
    case 2: dryv::export_all_kernel<2><<<blocks, 256, 0, st>>>(a); break;
    default: dryv::export_all_kernel<1><<<blocks, 256, 0, st>>>(a); break;
  }
  CU(cudaGetLastError());
  ctx->launches++;
  return DRYV_OK;
}

}  // namespace

extern "C" {

int dryv_recon_abi_version(void) { return DRYV_RECON_ABI_VERSION; }

size_t dryv_recon_frame_bytes(const dryv_pic_params* pp) {
  if (!pp) return 0;
  return (size_t)pp->pic_width_in_mbs * pp->pic_height_in_mbs * 384;
}

int dryv_recon_create(int device, dryv_recon_ctx** out) {
  if (!out) return DRYV_ERR_ARG;
  *out = nullptr;
  dryv_recon_ctx* ctx = new dryv_recon_ctx();
  ctx->device = device;
  auto bail = [&](int code) {
    delete ctx;
    return code;
  };
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return bail(DRYV_ERR_CUDA);
  if (cudaSetDevice(device) != cudaSuccess) return bail(DRYV_ERR_CUDA);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(DRYV_ERR_CUDA);
  if (prop.major != 10) {  // sm_100a SASS only: no PTX fallback, no other architecture
    fprintf(stderr, "dryv_recon: device %d is sm_%d%d, this library is built for sm_100a only\n", device, prop.major,
            prop.minor);
    return bail(DRYV_ERR_CUDA);
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->use_pdl = getenv("DRYV_NO_PDL") == nullptr;
  if (const char* g = getenv("DRYV_WAVE_GRID")) ctx->wave_grid_override = atoi(g);
  if (const char* g = getenv("DRYV_SPLIT")) ctx->use_split = atoi(g) != 0;
  bool ok = cudaStreamCreateWithFlags(&ctx->s_compute[0], cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->s_compute[1], cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < dryv_recon_ctx::kStages && ok; i++)
    ok = cudaEventCreateWithFlags(&ctx->e_h2d[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->e_kernel[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->e_d2h[i], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreate(&ctx->e_sub_begin) == cudaSuccess && cudaEventCreate(&ctx->e_sub_end) == cudaSuccess;
  for (int i = 0; i < dryv_recon_ctx::kPending && ok; i++)
    ok = cudaEventCreateWithFlags(&ctx->e_done[i], cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < dryv_recon_ctx::kTimedLaunches && ok; i++)
    ok = cudaEventCreate(&ctx->e_wave[i][0]) == cudaSuccess && cudaEventCreate(&ctx->e_wave[i][1]) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_tables, sizeof(DeviceTables)) == cudaSuccess &&
       cudaMallocHost(&ctx->h_tables, sizeof(DeviceTables)) == cudaSuccess &&
       cudaMalloc(&ctx->d_ticket, 2 * dryv_recon_ctx::kSets * sizeof(unsigned int)) == cudaSuccess &&
       cudaMalloc(&ctx->d_prof, 16 * sizeof(unsigned long long)) == cudaSuccess &&
       cudaMemset(ctx->d_prof, 0, 16 * sizeof(unsigned long long)) == cudaSuccess &&
       cudaMallocHost(&ctx->h_status, sizeof(int)) == cudaSuccess &&
       cudaMemset(ctx->d_ticket, 0, 2 * dryv_recon_ctx::kSets * sizeof(unsigned int)) == cudaSuccess;
  for (int i = 0; i < dryv_recon_ctx::kSets && ok; i++)
    ok = cudaEventCreateWithFlags(&ctx->ctl[i].done, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&ctx->db_done, cudaEventDisableTiming) == cudaSuccess;
  // shared memory, not L1, is what the row teams live on: ask for the largest carve-out
  int wave_carveout = cudaSharedmemCarveoutMaxShared;
  if (const char* g = getenv("DRYV_WAVE_CARVEOUT")) wave_carveout = atoi(g);  // development: percent of the L1 / shared array
  ok = ok && cudaFuncSetAttribute(dryv::recon_wavefront_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  wave_carveout) == cudaSuccess &&
       cudaFuncSetAttribute(dryv::recon_wavefront_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)sizeof(dryv::WaveCtaSmem)) == cudaSuccess;
  ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->wave_ctas_per_sm, dryv::recon_wavefront_kernel,
                                                           dryv::kWaveThreads, sizeof(dryv::WaveCtaSmem)) == cudaSuccess &&
       cudaFuncSetAttribute(dryv::recon_residual_add_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)sizeof(dryv::ResidCtaSmem)) == cudaSuccess &&
       cudaFuncSetAttribute(dryv::recon_residual_add_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared) == cudaSuccess &&
       cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->resid_ctas_per_sm, dryv::recon_residual_add_kernel,
                                                     dryv::kThreadsPerCta, sizeof(dryv::ResidCtaSmem)) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_ticket_c, dryv_recon_ctx::kSets * sizeof(unsigned int)) == cudaSuccess &&
       cudaFuncSetAttribute(dryv::recon_residual_fields_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)sizeof(dryv::ResidCtaSmem)) == cudaSuccess &&
       cudaFuncSetAttribute(dryv::recon_residual_fields_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared) == cudaSuccess &&
       cudaFuncSetAttribute(dryv::recon_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)sizeof(dryv::SplitCtaSmem)) == cudaSuccess &&
       cudaFuncSetAttribute(dryv::recon_predict_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared) == cudaSuccess &&
       cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->split_ctas_per_sm, dryv::recon_predict_kernel,
                                                     dryv::kSplitThreads, sizeof(dryv::SplitCtaSmem)) == cudaSuccess &&
       ctx->split_ctas_per_sm >= 1;
#if DRYV_CLUSTER > 1
  if (ok) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ctx->sm_count * ctx->wave_ctas_per_sm / dryv::kCluster * dryv::kCluster));
    cfg.blockDim = dim3(dryv::kWaveThreads);
    cfg.dynamicSmemBytes = sizeof(dryv::WaveCtaSmem);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = dryv::kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ok = cudaOccupancyMaxActiveClusters(&ctx->wave_clusters, dryv::recon_wavefront_kernel, &cfg) == cudaSuccess &&
         ctx->wave_clusters >= 1;
    if (getenv("DRYV_VERBOSE")) fprintf(stderr, "dryv: %d clusters of %d row teams resident (%d CTAs per SM by occupancy)\n",
                                         ctx->wave_clusters, dryv::kCluster, ctx->wave_ctas_per_sm);
  }
#endif
  if (!ok || ctx->wave_ctas_per_sm < 1 || ctx->resid_ctas_per_sm < 1) {
    dryv_recon_destroy(ctx);
    return DRYV_ERR_CUDA;
  }
  *out = ctx;
  return DRYV_OK;
}

void dryv_recon_destroy(dryv_recon_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < dryv_recon_ctx::kStages; i++) {
    if (ctx->e_h2d[i]) cudaEventDestroy(ctx->e_h2d[i]);
    if (ctx->e_kernel[i]) cudaEventDestroy(ctx->e_kernel[i]);
    if (ctx->e_d2h[i]) cudaEventDestroy(ctx->e_d2h[i]);
    if (ctx->d_in[i]) cudaFree(ctx->d_in[i]);
    if (ctx->d_out[i]) cudaFree(ctx->d_out[i]);
    if (ctx->d_exp[i]) cudaFree(ctx->d_exp[i]);
  }
  for (int i = 0; i < dryv_recon_ctx::kTimedLaunches; i++) {
    if (ctx->e_wave[i][0]) cudaEventDestroy(ctx->e_wave[i][0]);
    if (ctx->e_wave[i][1]) cudaEventDestroy(ctx->e_wave[i][1]);
  }
  for (int i = 0; i < dryv_recon_ctx::kPending; i++)
    if (ctx->e_done[i]) cudaEventDestroy(ctx->e_done[i]);
  if (ctx->e_sub_begin) cudaEventDestroy(ctx->e_sub_begin);
  if (ctx->e_sub_end) cudaEventDestroy(ctx->e_sub_end);
  for (int i = 0; i < 2; i++)
    if (ctx->s_compute[i]) cudaStreamDestroy(ctx->s_compute[i]);
  if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
  if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
  if (ctx->d_tables) cudaFree(ctx->d_tables);
  if (ctx->h_tables) cudaFreeHost(ctx->h_tables);
  for (int i = 0; i < dryv_recon_ctx::kSets; i++) {
    if (ctx->ctl[i].d_line) cudaFree(ctx->ctl[i].d_line);
    if (ctx->ctl[i].d_modes) cudaFree(ctx->ctl[i].d_modes);
    if (ctx->ctl[i].done) cudaEventDestroy(ctx->ctl[i].done);
  }
  if (ctx->d_ticket) cudaFree(ctx->d_ticket);
  if (ctx->d_ticket_c) cudaFree(ctx->d_ticket_c);
  for (int i = 0; i < dryv_recon_ctx::kSets; i++)
    if (ctx->ctl[i].d_resid) cudaFree(ctx->ctl[i].d_resid);
  if (ctx->d_db_ticket) cudaFree(ctx->d_db_ticket);
  if (ctx->d_db_line) cudaFree(ctx->d_db_line);
  if (ctx->db_done) cudaEventDestroy(ctx->db_done);
  for (auto& u : ctx->user_streams) cudaEventDestroy(u.done);
  if (ctx->d_prof) cudaFree(ctx->d_prof);
  if (ctx->h_status) cudaFreeHost(ctx->h_status);
  for (cudaEvent_t e : ctx->trace_ev) cudaEventDestroy(e);
  delete ctx;
}

const char* dryv_recon_last_error(dryv_recon_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int dryv_recon_alloc_pinned(size_t bytes, void** out) {
  if (!out || bytes == 0) return DRYV_ERR_ARG;
  *out = nullptr;
  return cudaMallocHost(out, bytes) == cudaSuccess ? DRYV_OK : DRYV_ERR_CUDA;
}
void dryv_recon_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

int dryv_recon_reconstruct_device(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const dryv_mb_soa* d_soa,
                                  uint32_t n_frames, uint8_t* d_out_yuv, void* cuda_stream) {
  if (!ctx) return DRYV_ERR_ARG;
  if (!pp_ok(pp) || !soa_ok(d_soa) || !d_out_yuv || n_frames == 0) return fail(ctx, DRYV_ERR_ARG, "bad argument");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->s_compute[0];
  int rc = ensure_tables(ctx, pp, s);
  if (rc != DRYV_OK) return rc;
  rc = launch_wavefront(ctx, pp, d_soa, n_frames, d_out_yuv, s);
  if (rc != DRYV_OK) return rc;
  if (cuda_stream) {
    const int nrc = note_user_stream(ctx, s);
    if (nrc != DRYV_OK) return nrc;
  }
  return DRYV_OK;
}

int dryv_recon_residual_add_device(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const dryv_mb_soa* d_soa,
                                   uint32_t n_frames, const uint8_t* d_pred_yuv, uint8_t* d_out_yuv,
                                   void* cuda_stream) {
  if (!ctx) return DRYV_ERR_ARG;
  if (!pp_ok(pp) || !soa_ok(d_soa) || !d_out_yuv || !d_pred_yuv || n_frames == 0)
    return fail(ctx, DRYV_ERR_ARG, "bad argument");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->s_compute[0];
  int rc = ensure_tables(ctx, pp, s);
  if (rc != DRYV_OK) return rc;
  KernelArgs a = make_args(ctx, pp, d_soa, n_frames, d_out_yuv);
  a.pred_in = d_pred_yuv;
  const size_t mbs = (size_t)n_frames * pp->pic_width_in_mbs * pp->pic_height_in_mbs;
  if (mbs >= (1ull << 31)) return fail(ctx, DRYV_ERR_ARG, "more than 2^31 macroblocks in one launch");
  const size_t n_groups = (size_t)n_frames * pp->pic_height_in_mbs * ((pp->pic_width_in_mbs + dryv::kGroupMbs - 1) / dryv::kGroupMbs);
  size_t want = (n_groups + dryv::kWarpsPerCta - 1) / dryv::kWarpsPerCta;
  size_t cap = (size_t)ctx->sm_count * ctx->resid_ctas_per_sm;
  int grid = (int)(want < cap ? want : cap);
  dryv::recon_residual_add_kernel<<<grid, dryv::kThreadsPerCta, sizeof(dryv::ResidCtaSmem), s>>>(a);
  CU(cudaGetLastError());
  ctx->launches++;
  if (cuda_stream) {
    const int nrc = note_user_stream(ctx, s);
    if (nrc != DRYV_OK) return nrc;
  }
  return DRYV_OK;
}

// Shared pipeline of the two host-buffer entry points. Chunks of pictures travel through kStages staging slots on
// three kinds of streams: H2D copies, kernels (two compute streams, one wavefront control block each, so the latency
// bound kernels of neighbouring chunks overlap) and D2H copies.
// the deblocking post-pass over n_frames pictures at d_yuv on stream st (arguments checked by the callers)
static int launch_deblock(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const dryv_mb_soa* d_soa, uint32_t n_frames,
                          int slice_alpha_c0_offset_div2, int slice_beta_offset_div2, uint8_t* d_yuv, cudaStream_t st) {
  // the hand-off buffer takes 192 bytes per macroblock: long batches go through it in chunks of pictures
  const size_t mbs_per_frame = (size_t)pp->pic_width_in_mbs * pp->pic_height_in_mbs;
  const size_t frame_bytes = mbs_per_frame * 384;
  constexpr size_t kDbMaxMbs = (size_t)512 << 20 >> 7;  // ~768 MB of words at most
  const uint32_t chunk = (uint32_t)std::max<size_t>(1, std::min<size_t>(n_frames, kDbMaxMbs / mbs_per_frame));
  if (!ctx->d_db_ticket) {
    CU(cudaMalloc(&ctx->d_db_ticket, sizeof(unsigned int)));
  }
  if (chunk * mbs_per_frame > ctx->db_line_cap) {
    CU(cudaDeviceSynchronize());
    if (ctx->d_db_line) cudaFree(ctx->d_db_line);
    ctx->d_db_line = nullptr;
    ctx->db_line_cap = 0;
    const size_t bytes = chunk * mbs_per_frame * dryv::kDbLineWords * sizeof(unsigned long long);
    CU(cudaMalloc(&ctx->d_db_line, bytes));
    CU(cudaMemset(ctx->d_db_line, 0, bytes));
    ctx->db_line_cap = chunk * mbs_per_frame;
    ctx->db_tag = 0;
  }
  if (ctx->db_used) CU(cudaStreamWaitEvent(st, ctx->db_done, 0));  // ticket and words serve one launch at a time
  for (uint32_t f0 = 0; f0 < n_frames; f0 += chunk) {
    const uint32_t nf = std::min(chunk, n_frames - f0);
    const size_t rows = (size_t)((nf + 1) / 2) * pp->pic_height_in_mbs;  // a warp walks the same row of two pictures
    if (++ctx->db_tag == 0) {  // tag wrap: stale words could match again
      CU(cudaMemsetAsync(ctx->d_db_line, 0, ctx->db_line_cap * dryv::kDbLineWords * sizeof(unsigned long long), st));
      ctx->db_tag = 1;
    }
    CU(cudaMemsetAsync(ctx->d_db_ticket, 0, sizeof(unsigned int), st));
    dryv::DeblockArgs a;
    memset(&a, 0, sizeof a);
    a.yuv = d_yuv + (size_t)f0 * frame_bytes;
    a.qp = d_soa->qp + (size_t)f0 * mbs_per_frame;
    a.t8x8 = d_soa->transform_size_8x8_flag + (size_t)f0 * mbs_per_frame;
    a.line = ctx->d_db_line;
    a.ticket = ctx->d_db_ticket;
    a.tag = ctx->db_tag;
    a.status = reinterpret_cast<int*>(ctx->d_ticket + 1);
    a.W = pp->pic_width_in_mbs;
    a.H = pp->pic_height_in_mbs;
    a.n_frames = (int)nf;
    a.cb_off = pp->chroma_qp_index_offset;
    a.cr_off = pp->second_chroma_qp_index_offset;
    a.off_a = 2 * slice_alpha_c0_offset_div2;
    a.off_b = 2 * slice_beta_offset_div2;
    const size_t want = (rows + dryv::kDbWarps - 1) / dryv::kDbWarps, cap = (size_t)ctx->sm_count * 4;
    dryv::deblock_wavefront_kernel<<<(unsigned)(want < cap ? want : cap), 32 * dryv::kDbWarps, 0, st>>>(a);
    CU(cudaGetLastError());
    ctx->launches++;
  }
  CU(cudaEventRecord(ctx->db_done, st));
  ctx->db_used = true;
  return DRYV_OK;
}

static int submit_impl(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const dryv_mb_soa* soa,
                       const dryv_mb_levels_compact* lv, uint32_t n_frames, uint8_t* out_yuv) {
  constexpr int kStages = dryv_recon_ctx::kStages;
  CU(cudaSetDevice(ctx->device));
  int rc = ensure_tables(ctx, pp, ctx->s_compute[0]);
  if (rc != DRYV_OK) return rc;
  const size_t n_mb = (size_t)pp->pic_width_in_mbs * pp->pic_height_in_mbs;
  const size_t out_per_frame = n_mb * 384;
  const size_t dense_per_frame = n_mb * (4 + 16 + 768);
  const bool exporting = ctx->surface_set;
  const dryv_surface surf = ctx->surface;
  if (exporting && ((size_t)surf.crop_left + surf.width > 16u * (size_t)pp->pic_width_in_mbs ||
                    (size_t)surf.crop_top + surf.height > 16u * (size_t)pp->pic_height_in_mbs))
    return fail(ctx, DRYV_ERR_ARG, "surface rectangle leaves the coded picture");
  const size_t host_per_frame = exporting ? surface_bytes(&surf) : out_per_frame;  // what out_yuv receives per picture
  // chunk = pictures per pipeline stage. Measured on a B200 / PCIe Gen5 box (64 x 1080p):
  //  dense levels (412 MB in, 201 MB out): the H2D copy is the floor (7.4 ms alone, 8.0 ms with the D2H running
  //    against it); ~72 MB of input per stage keeps the per-chunk kernels hidden under the copies while the pipeline
  //    fill and drain stay short;
  //  compact levels: the D2H copy of the pictures is the floor, so stages are sized by output bytes.
  // Pictures are spread evenly over the stages so there is no runt at the end.
  uint32_t n_chunks;
  if (lv) {
    const char* env = getenv("DRYV_CHUNK_OUT_MB");  // development knob
    const size_t chunk_bytes = (size_t)(env ? atoi(env) : 24) << 20;
    n_chunks = (uint32_t)((out_per_frame * n_frames + chunk_bytes - 1) / chunk_bytes);
  } else {
    const char* env = getenv("DRYV_CHUNK_MB");  // development knob
    const size_t chunk_bytes = (size_t)(env ? atoi(env) : 72) << 20;
    n_chunks = (uint32_t)((dense_per_frame * n_frames + chunk_bytes - 1) / chunk_bytes);
  }
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > n_frames) n_chunks = n_frames;
  const uint32_t chunk = (n_frames + n_chunks - 1) / n_chunks;
  // Measured dead ends on the compact path (64 x 1080p, 4.9 ms): a short first stage to start the D2H sooner, and 8
  // staging slots instead of 4 (no change): once the D2H runs, H2D + D2H together move ~75 GB/s over the link (the
  // same total as a bidirectional copy probe), so the floor is (in + out bytes) / 75 GB/s plus one kernel latency.
  std::vector<uint32_t> sched;
  for (uint32_t left = n_frames; left;) {
    const uint32_t nf = left < chunk ? left : chunk;
    sched.push_back(nf);
    left -= nf;
  }
  // staging slot: dense levels | pred_syntax | mb_type | t8x8 | chroma mode | qp | [offsets | compact stream]
  size_t stream_max = 0;
  if (lv) {
    const size_t total_mbs = n_mb * n_frames;
    if (lv->offset[0] > lv->offset[total_mbs]) return fail(ctx, DRYV_ERR_ARG, "compact level stream: offsets not monotone");
    uint32_t done = 0;
    for (uint32_t nf : sched) {
      const uint32_t o0 = lv->offset[(size_t)done * n_mb], o1 = lv->offset[(size_t)(done + nf) * n_mb];
      if (o1 < o0 || (o0 & 3u)) return fail(ctx, DRYV_ERR_ARG, "compact level stream: offsets not monotone / not 4-byte aligned");
      if ((size_t)(o1 - o0) > stream_max) stream_max = o1 - o0;
      done += nf;
    }
  }
  const size_t off_bytes = lv ? (((size_t)chunk * n_mb + 1) * 4 + 15) & ~(size_t)15 : 0;
  const size_t need_in = dense_per_frame * chunk + off_bytes + ((stream_max + 15) & ~(size_t)15);
  const size_t need_out = out_per_frame * chunk;
  const size_t need_exp = exporting ? host_per_frame * chunk : 0;
  if (need_exp > ctx->exp_cap) {
    CU(cudaDeviceSynchronize());
    for (int i = 0; i < kStages; i++) {
      if (ctx->d_exp[i]) cudaFree(ctx->d_exp[i]);
      ctx->d_exp[i] = nullptr;
    }
    ctx->exp_cap = 0;
    for (int i = 0; i < kStages; i++) CU(cudaMalloc(&ctx->d_exp[i], need_exp));
    ctx->exp_cap = need_exp;
  }
  if (need_in > ctx->in_cap || need_out > ctx->out_cap) {
    CU(cudaDeviceSynchronize());
    for (int i = 0; i < kStages; i++) {
      if (ctx->d_in[i]) cudaFree(ctx->d_in[i]);
      if (ctx->d_out[i]) cudaFree(ctx->d_out[i]);
      ctx->d_in[i] = ctx->d_out[i] = nullptr;
    }
    const size_t cap_in = need_in > ctx->in_cap ? need_in : ctx->in_cap, cap_out = need_out > ctx->out_cap ? need_out : ctx->out_cap;
    ctx->in_cap = ctx->out_cap = 0;
    for (int i = 0; i < kStages; i++) {
      CU(cudaMalloc(&ctx->d_in[i], cap_in));
      CU(cudaMalloc(&ctx->d_out[i], cap_out));
    }
    ctx->in_cap = cap_in;
    ctx->out_cap = cap_out;
  }
  uint32_t done = 0;
  ctx->sub_timed = false;
  const bool trace = getenv("DRYV_SUBMIT_TRACE") != nullptr;
  for (cudaEvent_t e : ctx->trace_ev) cudaEventDestroy(e);
  ctx->trace_ev.clear();
  auto mark = [&](cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ctx->trace_ev.push_back(e);
  };
  // queue depth: block on the oldest outstanding submit when the completion ring is full
  if (ctx->pending_count == dryv_recon_ctx::kPending) {
    CU(cudaEventSynchronize(ctx->e_done[ctx->pending_head]));
    ctx->pending_head = (ctx->pending_head + 1) % dryv_recon_ctx::kPending;
    ctx->pending_count--;
  }
  if (!ctx->sub_open) CU(cudaEventRecord(ctx->e_sub_begin, ctx->s_h2d));  // timing spans every submit queued before the wait
  ctx->sub_open = true;
  for (uint32_t i = 0; i < (uint32_t)sched.size(); i++) {
    const int slot = (int)(i % kStages);
    cudaStream_t sc = ctx->s_compute[i & 1];
    const uint32_t nf = sched[i];
    const size_t mb0 = (size_t)done * n_mb, cnt = (size_t)nf * n_mb;
    uint8_t* base = ctx->d_in[slot];
    dryv_mb_soa d;
    d.coeff = reinterpret_cast<const int16_t*>(base);
    d.pred_syntax = base + cnt * 768;
    d.mb_type = base + cnt * (768 + 16);
    d.transform_size_8x8_flag = base + cnt * (768 + 17);
    d.intra_chroma_pred_mode = base + cnt * (768 + 18);
    d.qp = base + cnt * (768 + 19);
    uint32_t* d_off = reinterpret_cast<uint32_t*>(base + dense_per_frame * chunk);
    uint8_t* d_str = base + dense_per_frame * chunk + off_bytes;
    // H2D: the slot's previous kernel must have consumed its inputs
    const bool reused = ctx->slot_used[slot];  // also true for the first chunks of a submit queued behind another one
    if (reused) CU(cudaStreamWaitEvent(ctx->s_h2d, ctx->e_kernel[slot], 0));
    uint32_t o0 = 0, o1 = 0;
    if (lv) {
      o0 = lv->offset[mb0];
      o1 = lv->offset[mb0 + cnt];
      CU(cudaMemcpyAsync(d_off, lv->offset + mb0, (cnt + 1) * 4, cudaMemcpyHostToDevice, ctx->s_h2d));
      if (o1 > o0) CU(cudaMemcpyAsync(d_str, lv->stream + o0, o1 - o0, cudaMemcpyHostToDevice, ctx->s_h2d));
    } else {
      CU(cudaMemcpyAsync(const_cast<int16_t*>(d.coeff), soa->coeff + mb0 * 384, cnt * 768, cudaMemcpyHostToDevice, ctx->s_h2d));
    }
    CU(cudaMemcpyAsync(const_cast<uint8_t*>(d.pred_syntax), soa->pred_syntax + mb0 * 16, cnt * 16, cudaMemcpyHostToDevice, ctx->s_h2d));
    CU(cudaMemcpyAsync(const_cast<uint8_t*>(d.mb_type), soa->mb_type + mb0, cnt, cudaMemcpyHostToDevice, ctx->s_h2d));
    CU(cudaMemcpyAsync(const_cast<uint8_t*>(d.transform_size_8x8_flag), soa->transform_size_8x8_flag + mb0, cnt, cudaMemcpyHostToDevice, ctx->s_h2d));
    CU(cudaMemcpyAsync(const_cast<uint8_t*>(d.intra_chroma_pred_mode), soa->intra_chroma_pred_mode + mb0, cnt, cudaMemcpyHostToDevice, ctx->s_h2d));
    CU(cudaMemcpyAsync(const_cast<uint8_t*>(d.qp), soa->qp + mb0, cnt, cudaMemcpyHostToDevice, ctx->s_h2d));
    CU(cudaEventRecord(ctx->e_h2d[slot], ctx->s_h2d));
    mark(ctx->s_h2d);
    // kernels: inputs landed, the slot's previous output has been copied out
    CU(cudaStreamWaitEvent(sc, ctx->e_h2d[slot], 0));
    if (reused) CU(cudaStreamWaitEvent(sc, ctx->e_d2h[slot], 0));
    ctx->slot_used[slot] = true;
    mark(sc);
    if (lv) {
      rc = launch_expand(ctx, d_off, d_str, o0, o1 - o0, cnt, const_cast<int16_t*>(d.coeff), sc);
      if (rc != DRYV_OK) return rc;
    }
    rc = launch_wavefront(ctx, pp, &d, nf, ctx->d_out[slot], sc);
    if (rc != DRYV_OK) return rc;
    if (ctx->deblock_on) {
      rc = launch_deblock(ctx, pp, &d, nf, ctx->db_alpha_div2, ctx->db_beta_div2, ctx->d_out[slot], sc);
      if (rc != DRYV_OK) return rc;
    }
    if (exporting) {
      rc = launch_export(ctx, pp, ctx->d_out[slot], nf, &surf, ctx->d_exp[slot], sc);
      if (rc != DRYV_OK) return rc;
    }
    CU(cudaEventRecord(ctx->e_kernel[slot], sc));
    mark(sc);
    // D2H
    CU(cudaStreamWaitEvent(ctx->s_d2h, ctx->e_kernel[slot], 0));
    CU(cudaMemcpyAsync(out_yuv + (size_t)done * host_per_frame, exporting ? ctx->d_exp[slot] : ctx->d_out[slot],
                       (size_t)nf * host_per_frame, cudaMemcpyDeviceToHost, ctx->s_d2h));
    CU(cudaEventRecord(ctx->e_d2h[slot], ctx->s_d2h));
    mark(ctx->s_d2h);
    done += nf;
  }
  CU(cudaEventRecord(ctx->e_sub_end, ctx->s_d2h));
  CU(cudaEventRecord(ctx->e_done[(ctx->pending_head + ctx->pending_count) % dryv_recon_ctx::kPending], ctx->s_d2h));
  ctx->pending_count++;
  ctx->sub_timed = true;
  return DRYV_OK;
}

int dryv_recon_submit(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const dryv_mb_soa* soa, uint32_t n_frames,
                      uint8_t* out_yuv) {
  if (!ctx) return DRYV_ERR_ARG;
  if (!pp_ok(pp) || !soa_ok(soa) || !out_yuv || n_frames == 0) return fail(ctx, DRYV_ERR_ARG, "bad argument");
  return submit_impl(ctx, pp, soa, nullptr, n_frames, out_yuv);
}

int dryv_recon_submit_compact(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const dryv_mb_soa* soa,
                              const dryv_mb_levels_compact* levels, uint32_t n_frames, uint8_t* out_yuv) {
  if (!ctx) return DRYV_ERR_ARG;
  const bool soa_fields = soa && soa->mb_type && soa->transform_size_8x8_flag && soa->intra_chroma_pred_mode && soa->qp &&
                          soa->pred_syntax;
  if (!pp_ok(pp) || !soa_fields || !levels || !levels->offset || !levels->stream || !out_yuv || n_frames == 0)
    return fail(ctx, DRYV_ERR_ARG, "bad argument");
  return submit_impl(ctx, pp, soa, levels, n_frames, out_yuv);
}

int dryv_recon_deblock_device(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const dryv_mb_soa* d_soa, uint32_t n_frames,
                              int slice_alpha_c0_offset_div2, int slice_beta_offset_div2, uint8_t* d_yuv, void* cuda_stream) {
  if (!ctx) return DRYV_ERR_ARG;
  if (!pp_ok(pp) || !d_soa || !d_soa->qp || !d_soa->transform_size_8x8_flag || !d_yuv || n_frames == 0 ||
      (reinterpret_cast<uintptr_t>(d_yuv) & 15u) || slice_alpha_c0_offset_div2 < -6 || slice_alpha_c0_offset_div2 > 6 ||
      slice_beta_offset_div2 < -6 || slice_beta_offset_div2 > 6)
    return fail(ctx, DRYV_ERR_ARG, "bad argument");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->s_compute[0];
  const int rc = launch_deblock(ctx, pp, d_soa, n_frames, slice_alpha_c0_offset_div2, slice_beta_offset_div2, d_yuv, st);
  if (rc != DRYV_OK) return rc;
  if (cuda_stream) {
    const int nrc = note_user_stream(ctx, st);
    if (nrc != DRYV_OK) return nrc;
  }
  return DRYV_OK;
}

int dryv_recon_set_deblock(dryv_recon_ctx* ctx, int enable, int slice_alpha_c0_offset_div2, int slice_beta_offset_div2) {
  if (!ctx) return DRYV_ERR_ARG;
  if (enable && (slice_alpha_c0_offset_div2 < -6 || slice_alpha_c0_offset_div2 > 6 || slice_beta_offset_div2 < -6 ||
                 slice_beta_offset_div2 > 6))
    return fail(ctx, DRYV_ERR_ARG, "filter offsets outside -6..6");
  ctx->deblock_on = enable != 0;
  ctx->db_alpha_div2 = slice_alpha_c0_offset_div2;
  ctx->db_beta_div2 = slice_beta_offset_div2;
  return DRYV_OK;
}

size_t dryv_recon_surface_bytes(const dryv_surface* s) { return surface_ok(s) ? surface_bytes(s) : 0; }

int dryv_recon_export_device(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const uint8_t* d_yuv, uint32_t n_frames,
                             const dryv_surface* s, uint8_t* d_out, void* cuda_stream) {
  if (!ctx) return DRYV_ERR_ARG;
  if (!pp_ok(pp) || !d_yuv || !d_out || n_frames == 0 || !surface_ok(s)) return fail(ctx, DRYV_ERR_ARG, "bad argument");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->s_compute[0];
  int rc = launch_export(ctx, pp, d_yuv, n_frames, s, d_out, st);
  if (rc != DRYV_OK) return rc;
  if (cuda_stream) {
    const int nrc = note_user_stream(ctx, st);
    if (nrc != DRYV_OK) return nrc;
  }
  return DRYV_OK;
}

int dryv_recon_set_surface(dryv_recon_ctx* ctx, const dryv_surface* s) {
  if (!ctx) return DRYV_ERR_ARG;
  if (!s) {
    ctx->surface_set = false;
    return DRYV_OK;
  }
  if (!surface_ok(s)) return fail(ctx, DRYV_ERR_ARG, "malformed surface");
  ctx->surface = *s;
  ctx->surface_set = true;
  return DRYV_OK;
}

int dryv_recon_expand_levels_device(dryv_recon_ctx* ctx, const dryv_mb_levels_compact* d_levels, size_t n_mbs,
                                    int16_t* d_coeff, void* cuda_stream) {
  if (!ctx) return DRYV_ERR_ARG;
  if (!d_levels || !d_levels->offset || !d_levels->stream || !d_coeff || n_mbs == 0 ||
      (reinterpret_cast<uintptr_t>(d_coeff) % 16) != 0)
    return fail(ctx, DRYV_ERR_ARG, "bad argument");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->s_compute[0];
  // the stream length is only known on the device here: bound it by the last offset
  uint32_t o_first = 0, o_last = 0;
  CU(cudaMemcpyAsync(&o_first, d_levels->offset, 4, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(&o_last, d_levels->offset + n_mbs, 4, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (o_last < o_first) return fail(ctx, DRYV_ERR_ARG, "compact level stream: offsets not monotone");
  // the kernel reads record headers as 32-bit words: a misaligned first record or stream pointer would fault (a sticky
  // CUDA error) instead of being reported like every other malformed stream
  if ((o_first & 3u) || (reinterpret_cast<uintptr_t>(d_levels->stream) & 3u))
    return fail(ctx, DRYV_ERR_ARG, "compact level stream: stream pointer and first offset must be 4-byte aligned");
  int rc = launch_expand(ctx, d_levels->offset, d_levels->stream + o_first, o_first, o_last - o_first, n_mbs, d_coeff, s);
  if (rc != DRYV_OK) return rc;
  if (cuda_stream) {
    const int nrc = note_user_stream(ctx, s);
    if (nrc != DRYV_OK) return nrc;
  }
  return DRYV_OK;
}

double dryv_recon_last_submit_ms(dryv_recon_ctx* ctx) {
  if (!ctx || !ctx->sub_timed) return -1.0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ctx->e_sub_begin, ctx->e_sub_end) != cudaSuccess) return -1.0;
  return (double)ms;
}

int dryv_recon_wait_oldest(dryv_recon_ctx* ctx) {
  if (!ctx) return DRYV_ERR_ARG;
  if (ctx->pending_count == 0) return DRYV_OK;
  CU(cudaSetDevice(ctx->device));
  CU(cudaEventSynchronize(ctx->e_done[ctx->pending_head]));
  ctx->pending_head = (ctx->pending_head + 1) % dryv_recon_ctx::kPending;
  ctx->pending_count--;
  return DRYV_OK;
}

int dryv_recon_wait(dryv_recon_ctx* ctx) {
  if (!ctx) return DRYV_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  ctx->pending_head = ctx->pending_count = 0;
  ctx->sub_open = false;
  CU(cudaStreamSynchronize(ctx->s_h2d));
  CU(cudaStreamSynchronize(ctx->s_compute[0]));
  CU(cudaStreamSynchronize(ctx->s_compute[1]));
  CU(cudaStreamSynchronize(ctx->s_d2h));
  {
    const int rc = sync_user_streams(ctx);
    if (rc != DRYV_OK) return rc;
  }
  for (int i = 0; i < dryv_recon_ctx::kSets; i++)  // launches on caller streams other than the last one
    if (ctx->ctl[i].used) CU(cudaEventSynchronize(ctx->ctl[i].done));
  if (ctx->db_used) CU(cudaEventSynchronize(ctx->db_done));
  if (!ctx->trace_ev.empty()) {
    for (size_t i = 0; i + 3 < ctx->trace_ev.size(); i += 4) {
      float t[4] = {0, 0, 0, 0};
      for (int k = 0; k < 4; k++) cudaEventElapsedTime(&t[k], ctx->e_sub_begin, ctx->trace_ev[i + k]);
      fprintf(stderr, "submit trace chunk %zu: h2d done %.3f  kernels start %.3f  done %.3f  d2h done %.3f ms\n", i / 4, t[0],
              t[1], t[2], t[3]);
    }
    for (cudaEvent_t e : ctx->trace_ev) cudaEventDestroy(e);
    ctx->trace_ev.clear();
  }
  CU(cudaMemcpy(ctx->h_status, ctx->d_ticket + 1, sizeof(int), cudaMemcpyDeviceToHost));
  const int st = *ctx->h_status;
  if (st != dryv::STATUS_OK) {
    CU(cudaMemset(ctx->d_ticket + 1, 0, sizeof(int)));
    if (st == dryv::STATUS_WATCHDOG) return fail(ctx, DRYV_ERR_WATCHDOG, "wavefront watchdog fired");
    return fail(ctx, DRYV_ERR_UNSUPPORTED, "unsupported macroblock syntax (mb_type > 24, chroma mode > 3 or qp > 51)");
  }
  return DRYV_OK;
}

int dryv_recon_write_yuv_file(const uint8_t* frame_yuv, size_t bytes, const char* path) {
  if (!frame_yuv || !path || bytes == 0) return DRYV_ERR_ARG;
  std::string p(path);
  size_t slash = p.find_last_of('/');
  if (slash != std::string::npos && slash > 0) {
    std::string dir = p.substr(0, slash);
    if (mkdir(dir.c_str(), 0777) != 0 && errno != EEXIST) return DRYV_ERR_ARG;
  }
  FILE* f = fopen(path, "wb");
  if (!f) return DRYV_ERR_ARG;
  size_t w = fwrite(frame_yuv, 1, bytes, f);
  fclose(f);
  return w == bytes ? DRYV_OK : DRYV_ERR_ARG;
}

uint64_t dryv_recon_launch_count(dryv_recon_ctx* ctx) { return ctx ? ctx->launches : 0; }

int dryv_recon_wavefront_times(dryv_recon_ctx* ctx, float* out_ms, int cap) {
  if (!ctx || !out_ms || cap <= 0) return DRYV_ERR_ARG;
  int n = (int)(ctx->wave_launches < (uint64_t)dryv_recon_ctx::kTimedLaunches ? ctx->wave_launches
                                                                              : (uint64_t)dryv_recon_ctx::kTimedLaunches);
  if (n > cap) n = cap;
  for (int i = 0; i < n; i++) {  // out_ms[0] = the most recent launch
    cudaEvent_t* ev = ctx->e_wave[(ctx->wave_launches - 1 - i) % dryv_recon_ctx::kTimedLaunches];
    if (cudaEventElapsedTime(&out_ms[i], ev[0], ev[1]) != cudaSuccess) return DRYV_ERR_CUDA;
  }
  return n;
}

size_t dryv_recon_device_tables(const dryv_pic_params* pp, void* out, size_t cap) {
  if (!pp) return 0;
  DeviceTables* t = new DeviceTables();
  dryv::build_device_tables(*pp, t);
  if (out) memcpy(out, t, cap < sizeof(DeviceTables) ? cap : sizeof(DeviceTables));
  delete t;
  return sizeof(DeviceTables);
}

#ifdef DRYV_TRACE
// development builds only: allocate / read the per-macroblock timeline of picture 0
int dryv_recon_debug_trace(dryv_recon_ctx* ctx, unsigned int* out, size_t mbs) {
  if (!ctx) return DRYV_ERR_ARG;
  if (!out) {
    if (ctx->d_trace) cudaFree(ctx->d_trace);
    ctx->d_trace = nullptr;
    if (cudaMalloc(&ctx->d_trace, mbs * 16) != cudaSuccess) return DRYV_ERR_CUDA;
    cudaMemset(ctx->d_trace, 0, mbs * 16);
    ctx->trace_mbs = mbs;
    return DRYV_OK;
  }
  if (mbs > ctx->trace_mbs) return DRYV_ERR_ARG;
  return cudaMemcpy(out, ctx->d_trace, mbs * 16, cudaMemcpyDeviceToHost) == cudaSuccess ? DRYV_OK : DRYV_ERR_CUDA;
}
#endif

#ifdef DRYV_STAGE_CLOCKS
// development builds only (not part of include/dryv_recon.h): read and reset the stage clocks
int dryv_recon_debug_clocks(dryv_recon_ctx* ctx, unsigned long long out[16]) {
  if (!ctx) return DRYV_ERR_ARG;
  if (cudaMemcpy(out, ctx->d_prof, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return DRYV_ERR_CUDA;
  cudaMemset(ctx->d_prof, 0, 16 * sizeof(unsigned long long));
  return DRYV_OK;
}
#endif

}  // extern "C"

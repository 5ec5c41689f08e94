// split_kernels.cuh — the "split" formulation of the reconstruction path (included by recon.cu inside namespace dryv, after
// the helpers it uses: wait_line_words, bulk_load, the mbarrier wrappers).
//
// The row teams of recon_wavefront_kernel bind three things into one CTA that have nothing to do with each other:
//   * the residual stage (dequantisation + inverse transforms): no dependency between macroblocks at all;
//   * luma prediction: the x + 2y wavefront (a macroblock needs its left, top and top-right neighbours);
//   * chroma prediction: an x + y wavefront of its own (left and top only), independent of luma.
// A team is two warps at 128 registers, so an SM holds 16 such warps, and both are serial instruction chains. Here the three
// parts are separate walkers:
//   recon_residual_fields_kernel  (recon.cu)  streams every macroblock's biased residual fields into KernelArgs::resid
//                                 (928 bytes per macroblock, the layout the predictors read from shared memory);
//   recon_predict_kernel          one WARP per macroblock row: luma walkers (Intra4x4 / 8x8 / 16x16 prediction + residual
//                                 add + clip, bottom line published as tagged words like the row teams do) and chroma
//                                 walkers (their own row tickets, chroma words of the line array), each fetching its
//                                 macroblock's residual tile with one bulk copy a macroblock ahead. 64 registers per
//                                 thread, ~2 KB of shared memory per walker: 32 walkers per SM instead of 8 teams.
// Reference behaviour: the same functions of recon_kernels.cuh (pred4x4.rs, pred8x8.rs, pred16x16.rs, trans_chroma.rs).

#ifndef DRYV_SPLIT_LUMA_WARPS
#define DRYV_SPLIT_LUMA_WARPS 6
#endif
#ifndef DRYV_SPLIT_CHROMA_WARPS
#define DRYV_SPLIT_CHROMA_WARPS 2
#endif
#ifndef DRYV_SPLIT_CTAS
#define DRYV_SPLIT_CTAS 4
#endif
// Sleep between the polls of a walker that waits for the row above (ns). With 32 walkers per SM most of them wait at any
// time, and a spinning warp takes issue slots from the working ones (first build: 360 of 800 warp-instructions per
// macroblock were polls).
#ifndef DRYV_SPLIT_LUMA_NS
#define DRYV_SPLIT_LUMA_NS 200u
#endif
#ifndef DRYV_SPLIT_CHROMA_NS
#define DRYV_SPLIT_CHROMA_NS 500u
#endif
#ifndef DRYV_SPLIT_START_NS
#define DRYV_SPLIT_START_NS 1000u
#endif
constexpr int kSplitLumaWarps = DRYV_SPLIT_LUMA_WARPS, kSplitChromaWarps = DRYV_SPLIT_CHROMA_WARPS;
constexpr int kSplitThreads = 32 * (kSplitLumaWarps + kSplitChromaWarps);

struct LumaWalkerSmem {
  alignas(16) uint16_t res[2][kResLumaTile];   // residual tiles: macroblock x and the one fetched ahead
  alignas(16) uint8_t luma[kLumaTileBytes];    // pixel tile
  alignas(16) uint8_t lcol[16];                // right-most column of the previous macroblock
  alignas(16) uint8_t e8[32];                  // filtered edge vector of the current Intra8x8 block
  alignas(16) uint8_t rows[32];                // Intra4x4 tap rows (MbSlot::rows)
  alignas(8) unsigned long long full[2];       // mbarriers: residual tile landed
  uint32_t pace;
};
struct ChromaWalkerSmem {
  alignas(16) uint16_t res[2][kResChromaMb];
  alignas(16) uint8_t chroma[2 * kChromaTileBytes];
  alignas(16) uint8_t ccol[16];
  alignas(8) unsigned long long full[2];
  uint32_t pace;
};
struct SplitCtaSmem {
  alignas(16) unsigned char tab[kTeamTableBytes];
  LumaWalkerSmem lw[kSplitLumaWarps];
  ChromaWalkerSmem cw[kSplitChromaWarps > 0 ? kSplitChromaWarps : 1];
};

// ---- luma walker: the pixel warp of the row teams, fed from global memory instead of a group slot --------------------
__device__ __forceinline__ void split_luma_walk(const KernelArgs& a, const DeviceTables& tab, LumaWalkerSmem& ws, int lane) {
  const int W = a.W, H = a.H, W1 = W - 1;
  const size_t n_mb = (size_t)W * H;
  const int strideY = W * 16;
  const uint32_t tag = a.tag;
  const unsigned total_rows = (unsigned)a.n_frames * (unsigned)H;
  uint32_t pace_addr = smem_u32(&ws.pace);
  const PixLane pl = make_pix_lane(lane);
  bool dead = false, unsupported = false;
  uint8_t* const fresh_dst = lane < 4 ? &ws.luma[12 + 4 * (5 + lane)] : nullptr;
  const uint8_t* const pub_src = lane < 4 ? &ws.luma[luma_at(4 * lane, 15)] : nullptr;
  uint8_t* const shift_dst = lane < 5 ? &ws.luma[12 + 4 * lane] : nullptr;
  const int my_first = luma_at(0, lane & 15), my_last = luma_at(15, lane & 15), my_left = luma_at(-1, lane & 15);
  const bool m_lane = lane == 8 || lane == 9;  // the two words of a mode record
  unsigned nf = 0;  // residual tiles fetched / consumed so far: tile k lives in stage k & 1, phase (k >> 1) & 1
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(a.ticket, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= total_rows) break;
    const int row = (int)(t / (unsigned)a.n_frames), frame = (int)(t % (unsigned)a.n_frames);
    const size_t mb_row0 = (size_t)frame * n_mb + (size_t)row * W;
    const bool availB = row > 0, publish = row + 1 < H;
    const unsigned long long* line_above = a.line + (mb_row0 - W) * kLineWords + lane;
    unsigned long long* line_mine = a.line + mb_row0 * kLineWords + lane;
    uint8_t* st_ptr = a.out + (size_t)frame * n_mb * 384 + (size_t)(16 * row + (lane & 15)) * strideY;
    const uint16_t* const res_row = a.resid + mb_row0 * kResidMbFields;
    const uint8_t* const mt_row = a.mb_type + mb_row0;
    const uint8_t* const t8_row = a.t8x8 + mb_row0;
    const unsigned long long* const modes_row = a.modes + mb_row0 * kModeWords + (lane & 1);
    // macroblock 0: residual tile, header bytes, mode record
    if (lane == 0) bulk_load(ws.res[nf & 1], res_row, kResLumaTile * 2, &ws.full[nf & 1]);
    uint32_t h_mt = __ldg(mt_row), h_t8 = __ldg(t8_row);
    unsigned long long mv = m_lane ? ld_relaxed_gpu_u64(modes_row) : 0ull;
    unsigned long long lv = 0;
    if (availB) {  // luma line 0 of the row above becomes "line x" of macroblock 0
      unsigned long long v0 = 0;
      if (lane < 4) v0 = ld_relaxed_gpu_u64(line_above);
      const uint32_t w = wait_line_words_ns(line_above, v0, lane < 4, tag, DRYV_SPLIT_START_NS, a.status, dead, pace_addr);
      if (lane < 4) *reinterpret_cast<uint32_t*>(fresh_dst) = w;
      __syncwarp();
      uint32_t sv = 0;
      if (lane < 5) sv = *reinterpret_cast<const uint32_t*>(shift_dst + 16);
      __syncwarp();
      if (lane < 5) *reinterpret_cast<uint32_t*>(shift_dst) = sv;
      if (lane < 4) {
        line_above += kLineWords;
        if (W1 > 0) lv = ld_relaxed_gpu_u64(line_above);
      }
    }
    for (int x = 0; x < W; x++) {
      uint32_t mt = h_mt;
      const uint32_t t8 = h_t8;
      const unsigned long long mvc = mv;
      const uint16_t* const res = ws.res[nf & 1];
      unsigned long long* const res_bar = &ws.full[nf & 1];
      const uint32_t res_phase = (nf >> 1) & 1u;
      nf++;
      if (x < W1) {  // the next macroblock's tile (its stage held macroblock x - 1: consumed), header bytes, mode record
        if (lane == 0) bulk_load(ws.res[nf & 1], res_row + (size_t)(x + 1) * kResidMbFields, kResLumaTile * 2, &ws.full[nf & 1]);
        h_mt = __ldg(mt_row + x + 1);
        h_t8 = __ldg(t8_row + x + 1);
        if (m_lane) mv = ld_relaxed_gpu_u64(modes_row + (size_t)(x + 1) * kModeWords);
      }
      if (mt > 24u) {  // I_PCM / inter: flagged, never decoded (slice/macroblock.rs:682-716)
        unsupported = true;
        mt = 24u;
      }
      const int mbcls = mt == 0u ? (t8 ? 1 : 0) : 2;
      const int mode16 = (int)((mt - 1u) & 3u);
      const bool availA = x > 0, availC = availB && x < W1, availD = availA && availB;
      uint32_t modes_lo = 0, modes_hi = 0;
      if (mbcls != 2) {
        // the pre-pass may still be running: check the record's tags, poll if it has not got here yet
        const uint32_t w = wait_line_words_ns(modes_row + (size_t)x * kModeWords, mvc, m_lane, tag, 500u, a.status, dead, pace_addr);
        modes_lo = __shfl_sync(0xffffffffu, w, 8);
        modes_hi = __shfl_sync(0xffffffffu, w, 9);
        if (mbcls == 0) {
          const int av = (availA ? 1 : 0) | (availB ? 2 : 0) | (availC ? 4 : 0) | (availD ? 8 : 0);
          if (lane < 16) ws.rows[lane] = (uint8_t)i4_tap_row(tab, lane, modes_lo, modes_hi, av);
        }
      }
      if (availC) {  // line x + 1 of the row above (top-right neighbour)
        const uint32_t w = wait_line_words_ns(line_above, lv, lane < 4, tag, DRYV_SPLIT_LUMA_NS, a.status, dead, pace_addr);
        if (lane < 4) *reinterpret_cast<uint32_t*>(fresh_dst) = w;
      }
      mbar_wait(res_bar, res_phase);
      __syncwarp();
      if (mbcls == 0) predict_i4x4(tab, ws.luma, res, pl, ws.rows);
      else if (mbcls == 1) predict_i8x8(tab, ws.luma, ws.e8, res, pl, lane, modes_lo, modes_hi, availA, availB, availC, availD);
      else predict_i16x16(ws.luma, ws.lcol, res, lane, mode16, availA, availB);
      // the row below waits for exactly this: publish before anything else, then fetch the line for the next macroblock
      if (lane < 4) {
        if (publish)
          st_relaxed_gpu_u64(line_mine, ((unsigned long long)tag << 32) | *reinterpret_cast<const uint32_t*>(pub_src));
        if (availB) {
          line_above += kLineWords;
          if (x + 2 <= W1) lv = ld_relaxed_gpu_u64(line_above);
        }
      }
      line_mine += kLineWords;
      if (lane < 16) {
        const uint4 pv = *reinterpret_cast<const uint4*>(&ws.luma[my_first]);
        __stcs(reinterpret_cast<uint4*>(st_ptr), pv);
        st_ptr += 16;
      }
      // carry: right-most column -> left-neighbour column, top-row slots shift by one macroblock
      const int cv = ws.luma[my_last];
      uint32_t sv = 0;
      if (shift_dst) sv = *reinterpret_cast<const uint32_t*>(shift_dst + 16);
      __syncwarp();
      if (lane < 16) {
        ws.luma[my_left] = (uint8_t)cv;
        ws.lcol[lane] = (uint8_t)cv;
      }
      if (shift_dst) *reinterpret_cast<uint32_t*>(shift_dst) = sv;
      __syncwarp();
    }
  }
  if (unsupported) atomicCAS(a.status, STATUS_OK, STATUS_UNSUPPORTED);
}

// ---- chroma walker: the chroma walk of the row teams' front warp ------------------------------------------------------
__device__ __forceinline__ void split_chroma_walk(const KernelArgs& a, ChromaWalkerSmem& ws, int lane) {
  const int W = a.W, H = a.H;
  const size_t n_mb = (size_t)W * H;
  const int strideC = W * 8;
  const uint32_t tag = a.tag;
  const unsigned total_rows = (unsigned)a.n_frames * (unsigned)H;
  uint32_t pace_addr = smem_u32(&ws.pace);
  bool dead = false, unsupported = false;
  // lanes 4..5 / 6..7 fetch and publish the two Cb / Cr words of a bottom line, lanes 8..9 shift the top-row slots,
  // lanes 16..23 / 24..31 store and carry one Cb / Cr pixel row each
  uint8_t* c_fresh = nullptr;
  const uint8_t* c_pub = nullptr;
  if (lane >= 4 && lane < 8) {
    const int pln = (lane - 4) >> 1, k = (lane - 4) & 1;
    c_fresh = &ws.chroma[pln * kChromaTileBytes + 4 + 4 * (1 + k)];
    c_pub = &ws.chroma[pln * kChromaTileBytes + chroma_at(4 * k, 7)];
  }
  uint8_t* const c_shift = (lane == 8 || lane == 9) ? &ws.chroma[(lane - 8) * kChromaTileBytes + 4] : nullptr;
  uint8_t* const c_tile = &ws.chroma[((lane >> 3) & 1) * kChromaTileBytes];
  const int c_first = chroma_at(0, lane & 7), c_last = chroma_at(7, lane & 7), c_left = chroma_at(-1, lane & 7);
  const bool c_lane = lane >= 4 && lane < 8;
  unsigned nf = 0;
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(a.ticket_c, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= total_rows) break;
    const int row = (int)(t / (unsigned)a.n_frames), frame = (int)(t % (unsigned)a.n_frames);
    const size_t mb_row0 = (size_t)frame * n_mb + (size_t)row * W;
    const bool availB = row > 0, publish = row + 1 < H;
    const unsigned long long* c_above = a.line + (mb_row0 - W) * kLineWords + lane;
    unsigned long long* c_mine = a.line + mb_row0 * kLineWords + lane;
    uint8_t* c_st = a.out + (size_t)frame * n_mb * 384 + n_mb * 256 + (size_t)((lane >> 3) & 1) * n_mb * 64 +
                    (size_t)(8 * row + (lane & 7)) * strideC;
    const uint16_t* const res_row = a.resid + mb_row0 * kResidMbFields + kResLumaTile;
    const uint8_t* const cm_row = a.chroma_mode + mb_row0;
    if (lane == 0) bulk_load(ws.res[nf & 1], res_row, kResChromaMb * 2, &ws.full[nf & 1]);
    uint32_t h_cm = __ldg(cm_row);
    unsigned long long lvc = 0;
    if (availB && c_lane) lvc = ld_relaxed_gpu_u64(c_above);
    for (int x = 0; x < W; x++) {
      uint32_t cm = h_cm;
      const uint16_t* const res = ws.res[nf & 1];
      unsigned long long* const res_bar = &ws.full[nf & 1];
      const uint32_t res_phase = (nf >> 1) & 1u;
      nf++;
      if (x + 1 < W) {
        if (lane == 0) bulk_load(ws.res[nf & 1], res_row + (size_t)(x + 1) * kResidMbFields, kResChromaMb * 2, &ws.full[nf & 1]);
        h_cm = __ldg(cm_row + x + 1);
      }
      if (cm > 3u) {
        unsupported = true;
        cm &= 3u;
      }
      const bool availA = x > 0;
      if (availB) {
        const uint32_t w = wait_line_words_ns(c_above, lvc, c_lane, tag, x == 0 ? DRYV_SPLIT_START_NS : DRYV_SPLIT_CHROMA_NS, a.status, dead, pace_addr);
        if (c_lane) {
          *reinterpret_cast<uint32_t*>(c_fresh) = w;
          c_above += kLineWords;
          if (x + 1 < W) lvc = ld_relaxed_gpu_u64(c_above);
        }
      }
      mbar_wait(res_bar, res_phase);
      __syncwarp();
      predict_chroma(ws.chroma, ws.ccol, res, lane, (int)cm, availA, availB, availA && availB);
      if (publish && c_lane)
        st_relaxed_gpu_u64(c_mine, ((unsigned long long)tag << 32) | *reinterpret_cast<const uint32_t*>(c_pub));
      c_mine += kLineWords;
      if (lane >= 16) {
        const uint2 pv = *reinterpret_cast<const uint2*>(&c_tile[c_first]);
        __stcs(reinterpret_cast<uint2*>(c_st), pv);
        c_st += 8;
      }
      {  // carry: right-most column -> left-neighbour column, top-row shift
        const int cv = c_tile[c_last];
        uint32_t sv = 0;
        if (c_shift) sv = *reinterpret_cast<const uint32_t*>(c_shift + 8);
        __syncwarp();
        if (lane >= 16) {
          c_tile[c_left] = (uint8_t)cv;
          ws.ccol[lane - 16] = (uint8_t)cv;
        }
        if (c_shift) *reinterpret_cast<uint32_t*>(c_shift) = sv;
        __syncwarp();
      }
    }
  }
  if (unsupported) atomicCAS(a.status, STATUS_OK, STATUS_UNSUPPORTED);
}

__global__ void __launch_bounds__(kSplitThreads, DRYV_SPLIT_CTAS) recon_predict_kernel(const KernelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SplitCtaSmem& cs = *reinterpret_cast<SplitCtaSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.tables);
    uint4* dst = reinterpret_cast<uint4*>(cs.tab);
    for (int i = threadIdx.x; i < (int)(kTeamTableBytes / 16); i += kSplitThreads) dst[i] = src[i];
    if (warp < kSplitLumaWarps) {
      LumaWalkerSmem& ws = cs.lw[warp];
      ws.rows[lane] = 0;  // bytes 16..31 stay zero: the look-ahead of the Intra4x4 loop reads up to byte 18
      if (lane == 0) {
        mbar_init(&ws.full[0], 1);
        mbar_init(&ws.full[1], 1);
        ws.pace = smem_u32(&ws.pace);
      }
    } else {
      ChromaWalkerSmem& ws = cs.cw[warp - kSplitLumaWarps];
      if (lane == 0) {
        mbar_init(&ws.full[0], 1);
        mbar_init(&ws.full[1], 1);
        ws.pace = smem_u32(&ws.pace);
      }
    }
  }
  __syncthreads();
  const DeviceTables& tab = *reinterpret_cast<const DeviceTables*>(cs.tab);
  if (warp < kSplitLumaWarps) split_luma_walk(a, tab, cs.lw[warp], lane);
  else split_chroma_walk(a, cs.cw[warp - kSplitLumaWarps], lane);
}

/* synth.h — seeded spec-legal syntax-buffer generator (see synth.c). */
#ifndef DRYV_SYNTH_H
#define DRYV_SYNTH_H
#include "../../include/dryv_recon.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dryv_synth_cfg {
  int32_t qp_base;            /* luma QP before jitter, e.g. 26 */
  int32_t qp_jitter;          /* per-MB uniform jitter in [-j, +j] */
  int32_t pct_i4x4;           /* percent Intra4x4 macroblocks */
  int32_t pct_i8x8;           /* percent Intra8x8; remainder is Intra16x16 */
  int32_t stress_pct;         /* percent MBs with the wide residual profile (clamps to 0/255 regularly) */
  int32_t zero_residual;      /* 1: all levels zero (prediction-only pictures) */
  int32_t qp_step_per_frame;  /* batch only: picture f uses qp_base + f * step (QP sweep) */
  int32_t standard_only;      /* 1: no Intra8x8 macroblock in column 0, the one place where the reference's luma deviates
                                 from the H.264 text, so that a conformant decoder agrees on the luma */
} dryv_synth_cfg;

/* One picture; all output arrays are for that picture (n_mb entries, pred_syntax 16/MB, coeff 384/MB). */
int dryv_synth_frame(const dryv_pic_params* pp, const dryv_synth_cfg* cfg, uint64_t seed, uint8_t* mb_type,
                     uint8_t* t8x8, uint8_t* chroma_mode, uint8_t* qp, uint8_t* pred_syntax, int16_t* coeff);

/* n_frames pictures, picture f seeded with seed0 + f, generated on n_threads host threads. */
int dryv_synth_batch(const dryv_pic_params* pp, const dryv_synth_cfg* cfg, uint64_t seed0, uint32_t n_frames,
                     uint8_t* mb_type, uint8_t* t8x8, uint8_t* chroma_mode, uint8_t* qp, uint8_t* pred_syntax,
                     int16_t* coeff, uint32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif

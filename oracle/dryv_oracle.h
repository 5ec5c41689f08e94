/*
 * dryv_oracle.h — CPU oracle for the reconstruction path (TEST INFRASTRUCTURE ONLY, see dryv_oracle.c).
 * Same data contract as include/dryv_recon.h so the tests feed both sides the same buffers.
 */
#ifndef DRYV_ORACLE_H
#define DRYV_ORACLE_H
#include "../include/dryv_recon.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { DRYV_ORACLE_ERR_ARG = -1, DRYV_ORACLE_ERR_UNSUPPORTED = -2 };

size_t dryv_oracle_frame_bytes(const dryv_pic_params* pp);
int dryv_oracle_reconstruct(const dryv_pic_params* pp, const dryv_mb_soa* soa, uint32_t n_frames,
                            uint8_t* out_yuv);
int dryv_oracle_reconstruct_mt(const dryv_pic_params* pp, const dryv_mb_soa* soa, uint32_t n_frames,
                               uint8_t* out_yuv, uint32_t n_threads);
int dryv_oracle_residual_add(const dryv_pic_params* pp, const dryv_mb_soa* soa, uint32_t n_frames,
                             const uint8_t* pred_yuv, uint8_t* out_yuv);
int dryv_oracle_block4x4(const dryv_pic_params* pp, int qp1y, int mode, const int16_t coeff_zz[16],
                         int32_t r_out[16]);
int dryv_oracle_block8x8(const dryv_pic_params* pp, int qp1y, const int16_t coeff_zz[64],
                         int32_t r_out[64]);
int dryv_oracle_write_yuv_file(const uint8_t* frame_yuv, size_t bytes, const char* path);

#ifdef __cplusplus
}
#endif
#endif

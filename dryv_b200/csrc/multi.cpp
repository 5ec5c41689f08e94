// multi.cpp — in-process multi-GPU dispatch of the host-buffer entry points: one dryv_recon_ctx and one
// host thread per device, independent IDR pictures dealt in contiguous blocks, outputs written to disjoint slices of the
// caller's buffer. Nothing is exchanged between devices (pictures never reference each other on this path: Frame::new
// starts every picture from zeroed planes, reference src/video/frame/mod.rs:29-46), so there is no collective and no NCCL.
#include <cuda_runtime.h>

#include <string>
#include <thread>
#include <vector>

#include "../../include/dryv_recon.h"

struct dryv_recon_multi {
  std::vector<int> devices;
  std::vector<dryv_recon_ctx*> ctx;
  std::string err;
};

namespace {
// contiguous block of pictures of part d (the first n % parts parts get one more): dryv_b200/shard.py frames_for_rank
void block_of(uint32_t n, int d, int parts, uint32_t* lo, uint32_t* hi) {
  const uint32_t base = n / (uint32_t)parts, extra = n % (uint32_t)parts;
  *lo = (uint32_t)d * base + ((uint32_t)d < extra ? (uint32_t)d : extra);
  *hi = *lo + base + ((uint32_t)d < extra ? 1u : 0u);
}
}  // namespace

extern "C" {

int dryv_recon_multi_create(const int* devices, int n_devices, dryv_recon_multi** out) {
  if (!out) return DRYV_ERR_ARG;
  *out = nullptr;
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) return DRYV_ERR_CUDA;
  dryv_recon_multi* m = new dryv_recon_multi();
  if (devices) {
    if (n_devices < 1) {
      delete m;
      return DRYV_ERR_ARG;
    }
    m->devices.assign(devices, devices + n_devices);
  } else {
    const int n = n_devices > 0 ? n_devices : visible;
    for (int d = 0; d < n; d++) m->devices.push_back(d);
  }
  for (int d : m->devices) {
    dryv_recon_ctx* c = nullptr;
    const int rc = (d >= 0 && d < visible) ? dryv_recon_create(d, &c) : DRYV_ERR_ARG;
    if (rc != DRYV_OK) {
      for (dryv_recon_ctx* k : m->ctx) dryv_recon_destroy(k);
      delete m;
      return rc;
    }
    m->ctx.push_back(c);
  }
  *out = m;
  return DRYV_OK;
}

void dryv_recon_multi_destroy(dryv_recon_multi* m) {
  if (!m) return;
  for (dryv_recon_ctx* c : m->ctx) dryv_recon_destroy(c);
  delete m;
}

int dryv_recon_multi_device_count(const dryv_recon_multi* m) { return m ? (int)m->ctx.size() : 0; }

const char* dryv_recon_multi_last_error(dryv_recon_multi* m) { return m ? m->err.c_str() : "null dispatcher"; }

// levels == NULL: dense levels in soa->coeff (dryv_recon_submit); else the compact stream (dryv_recon_submit_compact)
static int multi_run(dryv_recon_multi* m, const dryv_pic_params* pp, const dryv_mb_soa* soa, const dryv_mb_levels_compact* levels,
                     uint32_t n_frames, uint8_t* out_yuv, size_t out_bytes_per_frame) {
  if (!m || !pp || !soa || !out_yuv || n_frames == 0) return DRYV_ERR_ARG;
  const int parts = (int)m->ctx.size();
  const size_t n_mb = (size_t)pp->pic_width_in_mbs * pp->pic_height_in_mbs;
  std::vector<int> rc(parts, DRYV_OK);
  std::vector<std::string> msg(parts);
  std::vector<std::thread> th;
  for (int d = 0; d < parts; d++) {
    uint32_t lo, hi;
    block_of(n_frames, d, parts, &lo, &hi);
    if (hi == lo) continue;
    th.emplace_back([=, &rc, &msg]() {
      dryv_recon_ctx* c = m->ctx[d];
      const size_t mb0 = (size_t)lo * n_mb;
      dryv_mb_soa s = *soa;  // this device's slice of every array
      s.mb_type += mb0;
      s.transform_size_8x8_flag += mb0;
      s.intra_chroma_pred_mode += mb0;
      s.qp += mb0;
      s.pred_syntax += mb0 * 16;
      if (s.coeff) s.coeff += mb0 * DRYV_COEFFS_PER_MB;
      uint8_t* o = out_yuv + (size_t)lo * out_bytes_per_frame;
      int r;
      if (levels) {
        dryv_mb_levels_compact lv = *levels;
        lv.offset += mb0;  // records are located by absolute stream offsets: the stream pointer stays
        r = dryv_recon_submit_compact(c, pp, &s, &lv, hi - lo, o);
      } else {
        r = dryv_recon_submit(c, pp, &s, hi - lo, o);
      }
      if (r == DRYV_OK) r = dryv_recon_wait(c);
      rc[d] = r;
      if (r != DRYV_OK) msg[d] = dryv_recon_last_error(c);
    });
  }
  for (std::thread& t : th) t.join();
  for (int d = 0; d < parts; d++)
    if (rc[d] != DRYV_OK) {
      m->err = "device " + std::to_string(m->devices[d]) + ": " + msg[d];
      return rc[d];
    }
  return DRYV_OK;
}

int dryv_recon_multi_reconstruct(dryv_recon_multi* m, const dryv_pic_params* pp, const dryv_mb_soa* soa, uint32_t n_frames,
                                 uint8_t* out_yuv) {
  if (!soa || !soa->coeff) return DRYV_ERR_ARG;
  return multi_run(m, pp, soa, nullptr, n_frames, out_yuv, dryv_recon_frame_bytes(pp));
}

int dryv_recon_multi_reconstruct_compact(dryv_recon_multi* m, const dryv_pic_params* pp, const dryv_mb_soa* soa,
                                         const dryv_mb_levels_compact* levels, uint32_t n_frames, uint8_t* out_yuv) {
  if (!levels || !levels->offset || !levels->stream) return DRYV_ERR_ARG;
  return multi_run(m, pp, soa, levels, n_frames, out_yuv, dryv_recon_frame_bytes(pp));
}

}  // extern "C"

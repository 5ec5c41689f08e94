// Robustness harness for the CPU CABAC / MP4 host (dryv_b200/csrc/cabac_host.cpp): seeded mutations (byte flips,
// truncation, insertion, deletion) of a valid stream or MP4 file go through every entry point of include/dryv_cabac_host.h.
// Built with -fsanitize=address,undefined by tests/test_cabac_host.py: any out-of-bounds access, overflow or crash fails it;
// the calls may return any status. usage: cabac_fuzz <file> <seed> <iterations>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/dryv_cabac_host.h"

static unsigned long long rng_state;
static unsigned rnd() {
  rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
  return (unsigned)(rng_state >> 33);
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> good((size_t)n);
  if (fread(good.data(), 1, (size_t)n, f) != (size_t)n) return 2;
  fclose(f);
  rng_state = strtoull(argv[2], nullptr, 10);
  const int iters = atoi(argv[3]);
  long ok = 0, rejected = 0;
  for (int it = 0; it < iters; it++) {
    std::vector<uint8_t> d = good;
    switch (it == 0 ? 99 : rnd() % 4) {
      case 0:
        for (unsigned k = 1 + rnd() % 6; k; k--) d[rnd() % d.size()] = (uint8_t)rnd();
        break;
      case 1:
        d.resize(1 + rnd() % d.size());
        break;
      case 2: {
        const size_t at = rnd() % d.size();
        std::vector<uint8_t> ins(1 + rnd() % 40);
        for (auto& b : ins) b = (uint8_t)rnd();
        d.insert(d.begin() + (long)at, ins.begin(), ins.end());
        break;
      }
      case 3: {
        const size_t at = rnd() % d.size(), len = 1 + rnd() % 200;
        d.erase(d.begin() + (long)at, d.begin() + (long)(at + len < d.size() ? at + len : d.size()));
        break;
      }
      default:
        break;  // the first iteration: the unmodified input
    }
    if (d.empty()) continue;
    // an exact-size heap copy, so that a read past the end is seen by the sanitizer
    uint8_t* p = (uint8_t*)malloc(d.size());
    memcpy(p, d.data(), d.size());
    dryv_pic_params pp;
    uint32_t np = 0;
    dryv_surface sf;
    dryv_slice_info si;
    dryv_cabac_surface(p, d.size(), &sf);
    dryv_cabac_slice_info(p, d.size(), 0, &si);
    dryv_cabac_picture_params(p, d.size(), 1, &pp);
    if (dryv_cabac_scan(p, d.size(), &pp, &np) == 0 && np > 0 && np < 64 && pp.pic_width_in_mbs <= 64 && pp.pic_height_in_mbs <= 64) {
      const size_t mbs = (size_t)pp.pic_width_in_mbs * pp.pic_height_in_mbs * np;
      std::vector<uint8_t> mt(mbs), t8(mbs), cm(mbs), qp(mbs), ps(mbs * 16), cs(mbs * 1024 + 64);
      std::vector<int16_t> co(mbs * 384);
      std::vector<uint32_t> off(mbs + 1);
      const int r1 = dryv_cabac_parse(p, d.size(), &pp, np, mt.data(), t8.data(), cm.data(), qp.data(), ps.data(), co.data(), 2);
      const int r2 = dryv_cabac_parse_compact(p, d.size(), &pp, 0, np, mt.data(), t8.data(), cm.data(), qp.data(), ps.data(),
                                              off.data(), cs.data(), cs.size(), 1);
      if (it == 0 && (r1 != 0 || r2 != 0)) {
        printf("the unmodified input does not parse: %d %d\n", r1, r2);
        return 1;
      }
      (r1 == 0 ? ok : rejected)++;
    } else {
      if (it == 0) {
        printf("the unmodified input does not scan\n");
        return 1;
      }
      rejected++;
    }
    free(p);
  }
  printf("ok: %d inputs, %ld parsed, %ld rejected\n", iters, ok, rejected);
  return 0;
}

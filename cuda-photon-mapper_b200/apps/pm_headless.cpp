// pm_headless.cpp -- headless driver over the C-ABI (include/pmb200.h): the replacement for the reference's GLUT loop
// (simpleGLMain.cpp main / callbacksPBO.cpp display): initRandomNumbers once, then per frame emit + render, written as
// PPM (the uchar4 frame the reference would have put on screen) and optionally PFM (the float framebuffer).
//
//   pm_headless [--photons N] [--width W --height H] [--szimg S] [--frames F] [--time T] [--dt DT]
//               [--media 0|1] [--interp 0|1] [--knn K] [--out prefix] [--pfm] [--compare ref.ppm --eps E --threshold T]
//
// --compare mirrors the SDK sample's regression mode (simpleGL.cpp:354-368, PPMvsPPM with an epsilon/threshold pair):
// exit status 1 if more than `threshold` of the channel values differ from the stored image by more than `eps`.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pmb200.h"

static void die(pm_context *c, const char *what, int rc) {
  fprintf(stderr, "pm_headless: %s failed (%d): %s\n", what, rc, c ? pm_last_error(c) : "");
  exit(EXIT_FAILURE);
}
#define PMCK(call) do { int rc_ = (call); if (rc_ != PM_OK) die(ctx, #call, rc_); } while (0)

static bool write_ppm(const std::string &path, const std::vector<pm_uchar4> &img, int w, int h) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return false;
  fprintf(f, "P6\n%d %d\n255\n", w, h);
  for (int i = 0; i < w * h; i++) { unsigned char px[3] = {img[i].x, img[i].y, img[i].z}; fwrite(px, 1, 3, f); }
  fclose(f);
  return true;
}
static bool write_pfm(const std::string &path, const std::vector<float> &rgbf, int w, int h) {   // bottom-up rows, little endian
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return false;
  fprintf(f, "PF\n%d %d\n-1.0\n", w, h);
  for (int y = h - 1; y >= 0; y--)
    for (int x = 0; x < w; x++) fwrite(&rgbf[4 * ((size_t)y * w + x)], sizeof(float), 3, f);
  fclose(f);
  return true;
}
static bool read_ppm(const std::string &path, std::vector<unsigned char> &rgb, int &w, int &h) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  int maxv = 0;
  if (fscanf(f, "P6 %d %d %d", &w, &h, &maxv) != 3 || maxv != 255) { fclose(f); return false; }
  fgetc(f);
  rgb.resize((size_t)w * h * 3);
  bool ok = fread(rgb.data(), 1, rgb.size(), f) == rgb.size();
  fclose(f);
  return ok;
}

int main(int argc, char **argv) {
  long long photons = 10000;          // nrPhotons, PMK:29
  int w = 512, h = 512, szimg = 0, frames = 1, media = 0, interp = 0, knn = 0, pfm = 0;
  float t = 0.0f, dt = 0.01f, eps = 10.0f, threshold = 0.15f;
  std::string out = "frame", compare;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() -> const char * { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
    if (a == "--photons") photons = atoll(next());
    else if (a == "--width") w = atoi(next());
    else if (a == "--height") h = atoi(next());
    else if (a == "--szimg") szimg = atoi(next());
    else if (a == "--frames") frames = atoi(next());
    else if (a == "--time") t = (float)atof(next());
    else if (a == "--dt") dt = (float)atof(next());
    else if (a == "--media") media = atoi(next());
    else if (a == "--interp") interp = atoi(next());
    else if (a == "--knn") knn = atoi(next());
    else if (a == "--out") out = next();
    else if (a == "--pfm") pfm = 1;
    else if (a == "--compare") compare = next();
    else if (a == "--eps") eps = (float)atof(next());
    else if (a == "--threshold") threshold = (float)atof(next());
    else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
  }
  pm_context *ctx = nullptr;
  int rc = pm_create(&ctx, -1);
  if (rc != PM_OK) { fprintf(stderr, "pm_headless: %s\n", rc == PM_ERR_NO_DEVICE ? "no CUDA device (there is no CPU fallback)" : "pm_create failed"); return EXIT_FAILURE; }
  pm_scene sc;
  pm_scene_default(&sc);
  sc.sz_img = szimg > 0 ? szimg : h;
  sc.cam_ox = -(float)(w - sc.sz_img) / 2.0f;   // centre a non-square frame; 0 for the reference's square one
  PMCK(pm_set_scene(ctx, &sc));
  PMCK(pm_set_photon_count(ctx, photons));
  PMCK(pm_set_energy_scale(ctx, 10000.0f / (float)photons));   // the reference's weights are tuned for 10 000 photons
  PMCK(pm_init_random_table(ctx));                             // initRandomNumbers(), first frame only
  std::vector<pm_uchar4> rgba((size_t)w * h);
  std::vector<float> rgbf((size_t)w * h * 4);
  if (knn > 0) PMCK(pm_set_record_capacity(ctx, (int64_t)(2.6 * (double)photons) + 4096));
  for (int f = 0; f < frames; f++, t += dt) {
    if (knn > 0) {   // Mode B: k-nearest-photon estimate
      pm_uchar4 *d_rgba = nullptr; float *d_rgbf = nullptr;
      PMCK(pm_clear_map(ctx));
      PMCK(pm_trace(ctx, t, (media ? PM_TRACE_MEDIA : 0u) | PM_TRACE_RECORDS | PM_TRACE_NO_MAP));
      PMCK(pm_knn_build(ctx, PM_MAP_SURFACE));
      if (media) PMCK(pm_knn_build(ctx, PM_MAP_VOLUME));
      PMCK(pm_render_knn_host(ctx, t, media != 0, w, h, knn, INFINITY, 2.0e-4f * 10000.0f / (float)photons, 4.0e-3f * 10000.0f / (float)photons,
                              rgba.data(), rgbf.data()));
      (void)d_rgba; (void)d_rgbf;
    } else {         // Mode A: the reference's voxel-map estimator, display() order
      PMCK(pm_frame_host(ctx, t, true, interp != 0, media != 0, w, h, rgba.data(), rgbf.data()));
    }
    char name[64];
    snprintf(name, sizeof(name), "_%04d", f);
    std::string base = out + (frames > 1 ? name : "");
    if (!write_ppm(base + ".ppm", rgba, w, h)) { fprintf(stderr, "cannot write %s.ppm\n", base.c_str()); return EXIT_FAILURE; }
    if (pfm && !write_pfm(base + ".pfm", rgbf, w, h)) { fprintf(stderr, "cannot write %s.pfm\n", base.c_str()); return EXIT_FAILURE; }
    printf("frame %d t=%.3f -> %s.ppm (%lld kernels launched so far)\n", f, t, base.c_str(), (long long)pm_launch_count(ctx));
  }
  int status = EXIT_SUCCESS;
  if (!compare.empty()) {
    std::vector<unsigned char> ref; int rw = 0, rh = 0;
    if (!read_ppm(compare, ref, rw, rh) || rw != w || rh != h) { fprintf(stderr, "cannot read %s as a %dx%d P6 image\n", compare.c_str(), w, h); return EXIT_FAILURE; }
    size_t bad = 0;
    for (size_t i = 0; i < (size_t)w * h; i++) {
      const unsigned char px[3] = {rgba[i].x, rgba[i].y, rgba[i].z};
      for (int ch = 0; ch < 3; ch++) bad += fabsf((float)px[ch] - (float)ref[3 * i + ch]) > eps;
    }
    double frac = (double)bad / ((double)w * h * 3);
    printf("compare: %.4f%% of channel values differ by more than %.1f (threshold %.2f%%) -> %s\n", 100.0 * frac, eps, 100.0 * threshold,
           frac <= threshold ? "PASS" : "FAIL");
    if (frac > threshold) status = 1;
  }
  pm_destroy(ctx);
  return status;
}

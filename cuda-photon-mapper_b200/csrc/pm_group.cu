// pm_group.cu -- a single-process multi-GPU photon mapper behind the C-ABI (pm_group_*, include/pmb200.h).
//
// SURVEY.md 8(e) process model: "single process, one host thread per GPU".  The reference's caller is one host thread
// running display() (callbacksPBO.cpp:47-101: emit, render, every frame); pm_group_frame_host* is that call for n GPUs.
// The group owns n contexts (one per device), one non-blocking stream and one worker thread per context; a call posts the
// same command to every worker, which issues its rank's launches -- the ranks must progress concurrently because the
// accumulator exchange (pm_peer.cu) makes every rank wait for every other inside a kernel.
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

#include "pm_context.h"

namespace {

struct Worker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<int(pm_context *, int)> fn;
  std::atomic<uint64_t> posted{0}, done{0};
  std::atomic<bool> quit{false};
  int rc = PM_OK;   // first failing status since the last collect
};

}  // namespace

struct pm_group {
  int n = 0;
  std::vector<pm_context *> ctx;
  std::vector<cudaStream_t> streams;
  std::vector<std::unique_ptr<Worker>> workers;
  std::string err;
  int64_t next_ticket = 0;
  int64_t ctx_ticket[pm_context::kFrameRing][kMaxPeers];
  int64_t n_photons = 0;
  int64_t reserved_pixels = 0;
};

static void worker_main(pm_group *g, int rank) {
  Worker &w = *g->workers[rank];
  cudaSetDevice(g->ctx[rank]->device);
  uint64_t seen = 0;
  for (;;) {
    // a short spin keeps the hand-off latency of back-to-back frames in the microseconds; then sleep
    int spins = 0;
    while (w.posted.load(std::memory_order_acquire) == seen && !w.quit.load(std::memory_order_acquire)) {
      if (++spins < 20000) continue;
      std::unique_lock<std::mutex> lk(w.m);
      w.cv.wait(lk, [&] { return w.posted.load(std::memory_order_acquire) != seen || w.quit.load(std::memory_order_acquire); });
    }
    if (w.posted.load(std::memory_order_acquire) == seen) return;   // quit
    seen++;
    int rc = w.fn(g->ctx[rank], rank);
    if (rc != PM_OK && w.rc == PM_OK) w.rc = rc;
    {
      std::lock_guard<std::mutex> lk(w.m);
      w.done.store(seen, std::memory_order_release);
    }
    w.cv.notify_all();
  }
}

static void wait_all(pm_group *g) {
  for (auto &wp : g->workers) {
    Worker &w = *wp;
    int spins = 0;
    while (w.done.load(std::memory_order_acquire) != w.posted.load(std::memory_order_acquire)) {
      if (++spins < 20000) continue;
      std::unique_lock<std::mutex> lk(w.m);
      w.cv.wait(lk, [&] { return w.done.load(std::memory_order_acquire) == w.posted.load(std::memory_order_acquire); });
    }
  }
}

// hand `fn` to every worker (after its previous command has finished)
static void post_all(pm_group *g, std::function<int(pm_context *, int)> fn) {
  wait_all(g);
  for (auto &wp : g->workers) {
    Worker &w = *wp;
    {
      std::lock_guard<std::mutex> lk(w.m);
      w.fn = fn;
      w.posted.fetch_add(1, std::memory_order_release);
    }
    w.cv.notify_all();
  }
}

// first error any worker has reported since the last collect
static int collect(pm_group *g) {
  wait_all(g);
  int rc = PM_OK;
  for (int r = 0; r < g->n; r++) {
    Worker &w = *g->workers[r];
    if (w.rc != PM_OK && rc == PM_OK) {
      rc = w.rc;
      g->err = "rank " + std::to_string(r) + ": " + pm_last_error(g->ctx[r]);
    }
    w.rc = PM_OK;
  }
  return rc;
}

static int run_all(pm_group *g, std::function<int(pm_context *, int)> fn) {
  post_all(g, std::move(fn));
  return collect(g);
}

// Frame buffers are allocated by every rank BEFORE any rank starts the frame: an allocation synchronises with its device, and a
// rank that shares the device could already be spinning in the exchange kernel for the allocating rank (deadlock until the
// peer timeout; measured with three ranks on one GPU).
static int reserve(pm_group *g, int width, int height) {
  const int64_t pixels = (int64_t)width * height;
  if (pixels <= g->reserved_pixels) return PM_OK;
  int rc = run_all(g, [=](pm_context *c, int) { return pm_reserve_frame(c, width, height); });
  if (rc == PM_OK) g->reserved_pixels = pixels;
  return rc;
}

extern "C" {

int pm_group_create(pm_group **out, const int *devices, int n) {
  if (!out) return PM_ERR_ARG;
  *out = nullptr;
  if (!devices || n < 1 || n > kMaxPeers) return PM_ERR_ARG;
  pm_group *g = new pm_group();
  g->n = n;
  memset(g->ctx_ticket, 0, sizeof(g->ctx_ticket));
  int rc = PM_OK;
  for (int r = 0; r < n && rc == PM_OK; r++) {
    pm_context *c = nullptr;
    rc = pm_create(&c, devices[r]);
    if (rc != PM_OK) break;
    g->ctx.push_back(c);
    cudaStream_t st = nullptr;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { rc = PM_ERR_CUDA; break; }
    g->streams.push_back(st);
    pm_set_stream(c, st);
  }
  for (int r = 0; r < n && rc == PM_OK; r++) {
    rc = pm_peer_connect_local(g->ctx[r], r, n, g->ctx.data());
    if (rc != PM_OK) fprintf(stderr, "pmb200: pm_group_create: %s\n", pm_last_error(g->ctx[r]));
  }
  if (rc != PM_OK) {
    for (size_t r = 0; r < g->ctx.size(); r++) pm_destroy(g->ctx[r]);
    for (auto st : g->streams) cudaStreamDestroy(st);
    delete g;
    return rc;
  }
  for (int r = 0; r < n; r++) g->workers.emplace_back(new Worker());
  for (int r = 0; r < n; r++) g->workers[r]->th = std::thread(worker_main, g, r);
  *out = g;
  rc = pm_group_set_photon_count(g, 10000);   // nrPhotons, PMK:29
  if (rc != PM_OK) { pm_group_destroy(g); *out = nullptr; }
  return rc;
}

int pm_group_destroy(pm_group *g) {
  if (!g) return PM_ERR_ARG;
  wait_all(g);
  for (auto &wp : g->workers) {
    { std::lock_guard<std::mutex> lk(wp->m); wp->quit.store(true, std::memory_order_release); }
    wp->cv.notify_all();
  }
  for (auto &wp : g->workers) if (wp->th.joinable()) wp->th.join();
  for (int r = 0; r < g->n; r++) { cudaSetDevice(g->ctx[r]->device); cudaStreamSynchronize(g->streams[r]); }
  for (int r = 0; r < g->n; r++) pm_peer_disconnect(g->ctx[r]);
  for (int r = 0; r < g->n; r++) pm_destroy(g->ctx[r]);
  for (int r = 0; r < g->n; r++) cudaStreamDestroy(g->streams[r]);
  delete g;
  return PM_OK;
}

int pm_group_size(const pm_group *g) { return g ? g->n : 0; }
pm_context *pm_group_context(pm_group *g, int rank) { return (g && rank >= 0 && rank < g->n) ? g->ctx[rank] : nullptr; }
const char *pm_group_last_error(const pm_group *g) { return g ? g->err.c_str() : "null group"; }

int pm_group_set_scene(pm_group *g, const pm_scene *scene) {
  if (!g || !scene) return PM_ERR_ARG;
  pm_scene s = *scene;
  return run_all(g, [s](pm_context *c, int) { return pm_set_scene(c, &s); });
}

int pm_group_set_photon_count(pm_group *g, int64_t n_photons) {
  if (!g) return PM_ERR_ARG;
  const int n = g->n;
  int rc = run_all(g, [n_photons, n](pm_context *c, int r) {
    int rc = pm_set_photon_count(c, n_photons);
    if (rc != PM_OK) return rc;
    return pm_set_photon_range(c, n_photons * r / n, n_photons * (r + 1) / n);   // contiguous, disjoint, exhaustive
  });
  if (rc == PM_OK) g->n_photons = n_photons;
  return rc;
}

int pm_group_set_energy_scale(pm_group *g, float scale) {
  if (!g) return PM_ERR_ARG;
  return run_all(g, [scale](pm_context *c, int) { return pm_set_energy_scale(c, scale); });
}

int pm_group_init_random_table(pm_group *g) {
  if (!g) return PM_ERR_ARG;
  return run_all(g, [](pm_context *c, int) { return pm_init_random_table(c); });   // every rank: its own rows + rows 0..2
}

int pm_group_frame_host(pm_group *g, float t, bool emit, bool interp, bool media, int width, int height, pm_uchar4 *host_rgba, float *host_rgbf) {
  if (!g) return PM_ERR_ARG;
  if (width <= 0 || height <= 0) { g->err = "bad frame geometry"; return PM_ERR_ARG; }
  const int n = g->n;
  {
    int rc = reserve(g, width, height);
    if (rc != PM_OK) return rc;
  }
  return run_all(g, [=](pm_context *c, int r) {
    int rc = pm_set_row_band(c, (int)((int64_t)height * r / n), (int)((int64_t)height * (r + 1) / n));
    if (rc != PM_OK) return rc;
    return pm_frame_host(c, t, emit, interp, media, width, height, host_rgba, host_rgbf);
  });
}

int pm_group_frame_host_async(pm_group *g, float t, bool emit, bool interp, bool media, int width, int height, pm_uchar4 *host_rgba, int64_t *ticket) {
  if (!g || !ticket || !host_rgba) return PM_ERR_ARG;
  if (width <= 0 || height <= 0) { g->err = "bad frame geometry"; return PM_ERR_ARG; }
  const int n = g->n;
  {
    int rc = reserve(g, width, height);
    if (rc != PM_OK) return rc;
  }
  const int64_t tk = g->next_ticket++;
  int64_t *slot = g->ctx_ticket[tk % pm_context::kFrameRing];
  post_all(g, [=](pm_context *c, int r) {
    int rc = pm_set_row_band(c, (int)((int64_t)height * r / n), (int)((int64_t)height * (r + 1) / n));
    if (rc != PM_OK) return rc;
    return pm_frame_host_async(c, t, emit, interp, media, width, height, host_rgba, &slot[r]);
  });
  *ticket = tk;
  return PM_OK;
}

int pm_group_frame_wait(pm_group *g, int64_t ticket) {
  if (!g) return PM_ERR_ARG;
  if (!(ticket >= 0 && ticket < g->next_ticket && ticket + pm_context::kFrameRing >= g->next_ticket)) { g->err = "ticket is not one of the three most recent frames"; return PM_ERR_ARG; }
  int rc = collect(g);   // every rank has enqueued its part of every submitted frame
  if (rc != PM_OK) return rc;
  for (int r = 0; r < g->n; r++) {
    rc = pm_frame_wait(g->ctx[r], g->ctx_ticket[ticket % pm_context::kFrameRing][r]);
    if (rc != PM_OK) { g->err = "rank " + std::to_string(r) + ": " + pm_last_error(g->ctx[r]); return rc; }
  }
  return PM_OK;
}

}  // extern "C"

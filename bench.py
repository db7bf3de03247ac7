#!/usr/bin/env python
"""bench.py -- headline benchmark of pmb200 (the B200-native photon-mapping hot path).

Metric (BASELINE.json): ms/frame at 1920x1080 with a 16M-photon volumetric (participating-media) photon map
-- config 4, the configuration the north-star target is quoted on.  One "step" = one frame of the hot path:
clear map -> trace all photons (medium scattering on) -> [all-reduce of the exact accumulators when N > 1]
-> build map + gather tables -> eye rays + ray-march gather -> uchar4 frame + float4 framebuffer on rank 0.
The random-direction table is initialised once before the timed region, as the reference does (display():
callbacksPBO.cpp:55-58).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--photons P] [--width W --height H]

N > 1 is launched by torchrun, one rank per GPU (strong scaling: the photons and the screen rows are split).
--impl reference times the reference's own CPU implementation of the path (oracle/_ref, all host threads) on
a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ms/frame 1080p, 16M-photon volumetric map (trace + map build + ray-march render)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--photons", type=int, default=16777216)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip timing the reference CUDA kernels")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mode-b", action="store_true", help="skip the Mode B (k-NN photon map) sub-benchmarks")
    ap.add_argument("--mode", default="a", choices=["a", "b"],
                    help="a (default, headline): the reference's voxel-map estimator; b: k-NN photon map (records all-gathered, "
                         "tree built on every rank, row bands) -- for Mode B scaling runs")
    ap.add_argument("--knn", type=int, default=50, help="k of the Mode B estimate")
    ap.add_argument("--no-overlap", action="store_true",
                    help="Mode A: one stream, no frames in flight (the headline is a three-stage pipeline: exchange + map build of frame f on a second "
                         "stream and its render on a third, under the trace of frame f+1)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="Mode A, N > 1: how the accumulators are summed -- peer (default): our kernel pulls them over NVLink peer "
                         "memory inside pm_build_map (CUDA IPC between the ranks); nccl: dist.all_reduce (round 1's path)")
    ap.add_argument("--trace-sms", type=int, default=-1,
                    help="CTAs of the persistent trace kernel (pm_set_trace_sms); -1 (default): every SM at N = 1, all but 8 at N = 2, all but 16 at N >= 4, "
                         "which leaves room for the previous frames' exchange + map build and render on the other two streams")
    ap.add_argument("--no-extras", action="store_true", help="skip Mode B at N, config 5 and the single-GPU side benchmarks")
    ap.add_argument("--passes", type=int, default=1,
                    help="progressive photon mapping (BASELINE config 5): photon passes accumulated per frame, each with a fresh "
                         "direction table from the continuing MWC stream (Mode A)")
    return ap.parse_args()


def workload_config(a, extra=None):
    cfg = {"workload": "BASELINE config 4: default participating-media scene, %d photons, %dx%d, Mode A "
                       "(reference voxel-map estimator), media on, interpolate off" % (a.photons, a.width, a.height),
           "photons": a.photons, "passes": a.passes, "width": a.width, "height": a.height, "media": True, "interpolate": False,
           "rng": "MWC table (reference stream, jump-ahead)", "energy_scale": 10000.0 / a.photons,
           "cache": "inputs larger than L2: the 12 B/photon direction table (%.0f MB) is streamed every frame"
                    % (a.photons * 12 / 1e6)}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------------
# clocks: NVML sampled on a thread while the GPU is under load
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.thread, self.max_mhz = [], set(), False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own routines on the host cores (oracle/_ref), bounded sample
# ------------------------------------------------------------------------------------------------------
_TABLE_CACHE = {}


def cpu_reference_frame_ms(a, photon_div=1, row_div=1):
    """Returns (ms/frame, descriptor dict) of the reference's own routines on the host cores.  photon_div = row_div = 1:
    the whole frame of the bench configuration, nothing extrapolated.  Larger divisors (only used when the reference
    library is absent and the sequential port has to stand in): photons/photon_div traced + height/row_div rows rendered,
    each scaled back to the whole frame."""
    from oracle import oraclelib, refhost
    if not oraclelib.available():
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libpm_oracle.so"])
    orc = oraclelib.Oracle()
    n_s = max(a.photons // photon_div, 1000)
    rows = max(a.height // row_div, 1)
    if n_s not in _TABLE_CACHE:                   # the direction table is an input of the frame, generated once (display(): first frame only)
        _TABLE_CACHE.clear()
        _TABLE_CACHE[n_s] = orc.mwc_table(n_s)
    table, st = _TABLE_CACHE[n_s]
    if refhost.available():
        r = refhost.RefHost()
        try:   # all the host cores this process may run on, whatever OMP_NUM_THREADS a launcher (torchrun: 1) exported
            r.omp_set_threads(len(os.sched_getaffinity(0)))
        except AttributeError:
            r.omp_set_threads(os.cpu_count() or 1)
        cores = r.omp_threads()
        r.set_scene(sz_img=a.height)
        r.set_table(table); r.set_rng(*st); r.clear_grid()
        t0 = time.perf_counter()
        r.emit_omp(0, n_s, 0.0, False, True)
        t1 = time.perf_counter()
        # rows [0, rows) of the frame; the camera offset is applied like the product's cam_ox
        r.render_f32(a.width, rows, 0.0, False, True, ox=-(a.width - a.height) / 2.0, oy=(a.height - rows) / 2.0, omp=True)
        t2 = time.perf_counter()
        kind = "reference"
    else:
        sc = orc.default_scene(sz_img=a.height)
        sc.cam_ox = -(a.width - a.height) / 2.0
        cores = 1
        t0 = time.perf_counter()
        grid, _, _ = orc.emit(sc, table, 0, n_s, 0.0, True, rng=st)
        t1 = time.perf_counter()
        y0 = (a.height - rows) // 2
        orc.render(sc, grid, a.width, a.height, 0.0, False, True, y0=y0, y1=y0 + rows, want_u8=False)
        t2 = time.perf_counter()
        kind = "port"
    emit_ms = (t1 - t0) * 1e3 * (a.photons / n_s)
    render_ms = (t2 - t1) * 1e3 * (a.height / rows)
    what = ("reference routines (photonMappingKernel.cu:1-1521) as host C++ with OpenMP" if kind == "reference"
            else "sequential oracle port (oracle/_ref absent)")
    if n_s == a.photons and rows == a.height:
        sample = "the whole frame: %d photons traced + %d rows rendered, media on; %s" % (n_s, rows, what)
    else:
        sample = "%d of %d photons traced (x%.0f) + %d of %d rows rendered (x%.0f), media on; %s" % (
            n_s, a.photons, a.photons / n_s, rows, a.height, a.height / rows, what)
    return emit_ms + render_ms, {"kind": kind, "cores": cores, "sample": sample, "emit_ms": emit_ms, "render_ms": render_ms,
                                 "whole_frame": n_s == a.photons and rows == a.height}


def run_reference_arm(a):
    """bench.py --impl reference: the reference's own CPU implementation of the path, every host thread, on the SAME
    configuration (every step traces all the photons and renders all the rows: nothing is sampled or extrapolated)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refhost
    div = (1, 1) if refhost.available() else (64, 16)     # the single-threaded port cannot run 16M photons per step in minutes
    vals, walls = [], []
    desc = None
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        v, desc = cpu_reference_frame_ms(a, *div)
        if i >= a.warmup:
            vals.append(v); walls.append((time.perf_counter() - t0) * 1e3)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "ms", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": v, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(a),
            "note": "value = emit + render of one whole frame of the configuration (inputs resident in host memory, table generated "
                    "once); step wall time incl. clearing the grid: %.1f ms" % float(np.mean(walls)),
            "same_config": bool(desc["whole_frame"]),
            "cpu_baseline": {"value": v, "unit": "ms", "cores": desc["cores"], "kind": desc["kind"], "sample": desc["sample"]},
            "e2e": {"value": v, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# the reference CUDA kernels on this GPU (B-CUDA): reported beside our number, not a target
# ------------------------------------------------------------------------------------------------------
def time_reference_cuda(a, table_host, dev_rgba):
    path = os.path.join(ROOT, "oracle", "_ref", "libpmref_cuda_%d.so" % a.photons)
    if not os.path.exists(path):
        return {"unavailable": "oracle/_ref/libpmref_cuda_%d.so not built" % a.photons}
    L = C.CDLL(path)
    L.refcu_time_frames.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(C.c_float), C.POINTER(C.c_float)]
    assert L.refcu_capacity() == a.photons
    assert L.refcu_set_table(table_host.ctypes.data_as(C.c_void_p), a.photons) == 0
    assert L.refcu_set_szimg(a.height) == 0
    e, r = C.c_float(), C.c_float()
    rc = L.refcu_time_frames(C.c_void_p(dev_rgba.data_ptr()), a.width, a.height, 0.0, 0, 1, 1, 3, C.byref(e), C.byref(r))
    if rc != 0:
        return {"unavailable": "refcu_time_frames failed"}
    return {"ms_per_frame": e.value + r.value, "emit_ms": e.value, "render_ms": r.value, "steps": 3, "warmup": 1,
            "what": "photonMappingKernel.cu recompiled for sm_100a (nrPhotons=%d, szImg=%d), its own launchers, CUDA events"
                    % (a.photons, a.height)}


def mode_b_numbers(pmb200, torch, a, device):
    """Mode B (k-NN photon map) sub-benchmarks on one GPU: BASELINE configs 3 and 4.  Reported beside the headline."""
    out = {}
    W, H = a.width, a.height

    def ev(fn, reps, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for name, n, k, media in (("config3_surface_4M_k100", 4194304, 100, False), ("config4_volumetric_%dM_k50" % (a.photons >> 20), a.photons, 50, True)):
        m = pmb200.PhotonMapper(device=device, n_photons=n)
        m.set_stream(torch.cuda.current_stream().cuda_stream)
        sc = pmb200.default_scene(sz_img=H); sc.cam_ox = -(W - H) / 2.0
        m.set_scene(sc)
        m.init_random_numbers()
        m.set_record_capacity(int(2.6 * n) + 4096)

        def trace():
            m.clear_map(); m.trace(0.0, media=media, records=True, no_map=True)
        r = {"photons": n, "k": k, "media": media, "trace_with_records_ms": ev(trace, 3)}
        r["build_surface_ms"] = ev(lambda: m.knn_build(0), 3)
        n_rec = m.record_buffers(0)[3]
        r["surface_points"] = m.knn_size(0)[0]
        build_ms, build_pts = r["build_surface_ms"], n_rec
        if media:
            r["build_volume_ms"] = ev(lambda: m.knn_build(1), 3)
            r["volume_points"] = m.knn_size(1)[0]
            build_ms += r["build_volume_ms"]; build_pts += r["volume_points"]
        # algorithmic bytes of the build per sorted record: 16 read + 8 key/index written, 4 passes x (4 histogram
        # read + 8 read + 8 written), 4 + 16 + 16 for the permuted rows, boxes negligible
        r["build_GBps_algorithmic"] = build_pts * (24 + 4 * 20 + 36) / (build_ms * 1e-3) / 1e9
        rgba = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
        rgbf = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        ws, wv = 2.0e-4 * 10000.0 / n, 4.0e-3 * 10000.0 / n
        r["render_knn_ms"] = ev(lambda: m.render_knn(W, H, 0.0, media, k, float("inf"), ws, wv, rgba=rgba, rgbf=rgbf), 2)
        r["queries"] = W * H * (11 if media else 1)
        r["gather_queries_per_s"] = r["queries"] / (r["render_knn_ms"] * 1e-3)
        r["ms_per_frame"] = r["trace_with_records_ms"] + build_ms + r["render_knn_ms"]
        out[name] = r
        m.close()
        del rgba, rgbf
        torch.cuda.empty_cache()
    return out


def config2_numbers(pmb200, torch, device):
    """BASELINE config 2: default participating-media scene, 1M photons, 1024x1024, one B200, ours vs the reference CUDA kernel."""
    n, W, H = 1048576, 1024, 1024
    m = pmb200.PhotonMapper(device=device, n_photons=n)
    m.set_stream(torch.cuda.current_stream().cuda_stream)
    m.set_scene(pmb200.default_scene(sz_img=H))
    m.set_energy_scale(10000.0 / n)
    m.init_random_numbers()
    rgba = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    rgbf = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")

    def frame():
        m.clear_map(); m.trace(0.0, media=True); m.build_map()
        m.render_device(W, H, 0.0, False, True, rgba=rgba, rgbf=rgbf)
    for _ in range(5):
        frame()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        frame()
    e1.record(); torch.cuda.synchronize()
    out = {"workload": "default media scene, 1048576 photons, 1024x1024, Mode A", "ms_per_frame": e0.elapsed_time(e1) / 50}
    class A: pass
    a2 = A(); a2.photons, a2.width, a2.height = n, W, H
    out["reference_cuda_kernel"] = time_reference_cuda(a2, m.get_random_table(), rgba)
    m.close()
    return out


def _events(torch, n):
    return [torch.cuda.Event(enable_timing=True) for _ in range(n)]


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
        return

    import torch
    import torch.distributed as dist
    import pmb200
    from pmb200 import dist as pmdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pmb200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, NP = a.width, a.height, a.photons

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        tt = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    m = pmb200.PhotonMapper(device=local, n_photons=NP)
    main_stream = torch.cuda.current_stream()
    m.set_stream(main_stream.cuda_stream)
    scene = pmb200.default_scene(sz_img=H)
    scene.cam_ox = -(W - H) / 2.0
    m.set_scene(scene)
    m.set_energy_scale(10000.0 / NP / a.passes)
    first, last = pmdist.photon_shard(NP, rank, world)
    m.set_photon_range(first, last)
    m.init_random_numbers()                       # once, outside the timed region (callbacksPBO.cpp:55-58)
    y0, y1 = pmdist.row_band_uneven(H, rank, world)
    rows = y1 - y0
    m.set_row_band(y0, y1)
    # measured (gpurun, B200, three-stage frames: trace | exchange + map build | render + barrier): N = 2 is fastest with 140 of 148
    # trace CTAs (0.330 ms against 0.406 with all 148), N = 8 with 132 (0.125; 124: 0.128, 116: 0.135); N = 4 keeps 132 -- the shorter
    # the trace, the more of the frame is the chain of small kernels behind it, and that chain needs SMs of its own to keep up
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    trace_sms = a.trace_sms if a.trace_sms >= 0 else (0 if world == 1 else sms - 8 if world == 2 else sms - 16)
    m.set_trace_sms(trace_sms)

    # N > 1: the ranks' exchange blocks are mapped into each other (CUDA IPC) and the frame lives on rank 0, every rank
    # rendering its row band straight into it over NVLink
    peers = False
    if world > 1 and a.exchange == "peer":
        ok = torch.ones(1, device="cuda")
        try:
            pmdist.connect_peers(m)
        except pmb200.PmError as ex:
            print("rank %d: peer connect failed (%s), falling back to NCCL" % (rank, ex), file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        peers = bool(ok.item() > 0)
        if not peers:
            m.peer_disconnect()
    if peers:
        # the reference's output, the uchar4 frame, is assembled on rank 0 (7/8 of 8.3 MB inbound at 8 GPUs); the headless float4
        # framebuffer stays distributed, every rank keeping its band (assembling its 33 MB on one GPU saturated rank 0's NVLink
        # ingress: 40 us per frame at 8 GPUs)
        rgba_ptr, rgba_opened = pmdist.shared_frame(m, W * H * 4)
        rgba = pmdist.device_tensor(rgba_ptr, W * H * 4, "|u1").view(H, W, 4) if rank == 0 else rgba_ptr
        rgbf = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    else:
        rgba = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
        rgbf = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    acc_n = m.accumulators()[1]

    # ------------------------------------------------------------------------------------------------------
    # Mode A steps
    # ------------------------------------------------------------------------------------------------------
    def trace_passes():
        m.clear_map()
        m.trace(0.0, media=True)
        for _ in range(a.passes - 1):             # progressive: further passes accumulate into the same exact accumulators
            m.init_random_numbers()
            m.trace(0.0, media=True)

    def exchange_build_render(e=None):
        if world > 1 and not peers:
            pmdist.allreduce_accumulators(pmdist.device_tensor(m.accumulators()[0], acc_n, "<i8"))   # round 1's path
        if e: e[2].record()
        m.build_map()                             # peers: the exchange happens in here (peer_reduce_kernel)
        if e: e[3].record()
        m.render_device(W, H, 0.0, False, True, rgba=rgba, rgbf=rgbf, y0=y0, y1=y1)
        if peers:
            m.peer_barrier()                      # every rank's band has landed in rank 0's frame buffers
        elif world > 1:
            pmdist.gather_frame(rgba, y0, y1)

    def step_serial(e=None):
        """One frame on one stream, nothing in flight (--no-overlap, and the instrumented pass that splits the frame into stages)."""
        if e: e[0].record()
        trace_passes()
        if e: e[1].record()
        exchange_build_render(e)
        if e: e[4].record()

    side = torch.cuda.Stream()
    ev_traced = torch.cuda.Event()

    def step_pipelined(e=None):
        """Frames in flight: clear + trace on the main stream, exchange + map build on a second, render (+ barrier) on a third."""
        if a.passes == 1 and (peers or world == 1):
            m.frame_device(W, H, rgba=rgba, rgbf=rgbf, t=0.0, emit=True, interp=False, media=True)     # the library's own pipeline
            return
        trace_passes()
        ev_traced.record(main_stream)
        side.wait_event(ev_traced)
        m.set_stream(side.cuda_stream)
        with torch.cuda.stream(side):
            exchange_build_render()
        m.set_stream(main_stream.cuda_stream)

    def drain():
        m.sync()
        main_stream.wait_stream(side)

    # ------------------------------------------------------------------------------------------------------
    # Mode B step (records all-gathered with NCCL, trees built on every rank, interleaved rows)
    # ------------------------------------------------------------------------------------------------------
    side_b = torch.cuda.Stream() if world > 1 else None
    keep = {}

    def step_b(e=None):
        if e: e[0].record()
        m.clear_map()
        m.trace(0.0, media=True, records=True, no_map=True)
        if e: e[1].record()
        sp = pmdist.allgather_records(*[m.record_buffers(0)[i] for i in (0, 1, 3)])
        if e: e[2].record()
        if world > 1:                             # the volume records travel while the surface tree is being built
            ev_traced.record(main_stream)
            side_b.wait_event(ev_traced)
            with torch.cuda.stream(side_b):
                vp = pmdist.allgather_records(*[m.record_buffers(1)[i] for i in (0, 1, 3)])
            m.knn_build_points(0, sp[0], sp[1], sp[2], records=True)
            main_stream.wait_stream(side_b)
        else:
            vp = pmdist.allgather_records(*[m.record_buffers(1)[i] for i in (0, 1, 3)])
            m.knn_build_points(0, sp[0], sp[1], sp[2], records=True)
        m.knn_build_points(1, vp[0], vp[1], vp[2], records=True)
        if e: e[3].record()
        # the k-NN gather cost varies strongly over the image: rank r renders rows r, r+N, r+2N, ... and the frames are
        # summed (all other rows are zero, so the sum is exact)
        fb_u8, fb_f32 = keep["fb"]
        if world > 1:
            fb_u8.zero_(); fb_f32.zero_()
        m.render_knn(W, H, 0.0, True, a.knn, float("inf"), 2.0e-4 * 10000.0 / NP, 4.0e-3 * 10000.0 / NP, rgba=fb_u8, rgbf=fb_f32,
                     y0=rank, y1=H, y_step=world)
        if world > 1:
            dist.all_reduce(fb_f32)
            dist.all_reduce(fb_u8)
        if e: e[4].record()
        keep["rec"] = (sp, vp)                    # the maps reference the gathered arrays

    def time_steps(fn, steps, warm, events=None, after=None):
        for _ in range(warm):
            fn()
        if after: after()
        barrier()
        t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_beg.record()
        for i in range(steps):
            fn(events[i]) if events else fn()
        if after: after()                         # the last frame's second half is inside the timed span
        t_end.record()
        barrier()
        return max_over_ranks(t_beg.elapsed_time(t_end)) / steps

    if a.mode == "b":
        m.set_record_capacity(int(2.6 * (last - first)) + 4096)
        keep["fb"] = (torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda"), torch.zeros((H, W, 4), dtype=torch.float32, device="cuda"))
        step, after = step_b, None
    elif a.no_overlap:
        step, after = step_serial, None
    else:
        step, after = step_pipelined, drain

    # ---- the headline: K steps, nothing but the frames in the timed region ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler: sampler.start()
    launches0 = m.launch_count()
    ms_per_step = time_steps(step, a.steps, max(a.warmup, 3), after=after)
    launches = (m.launch_count() - launches0) * a.steps // (a.steps + max(a.warmup, 3))

    # ---- instrumented pass: the same frames on one stream with CUDA events between the stages and around every kernel ----
    n_inst = max(3, min(a.steps, 20))
    ev = [_events(torch, 5) for _ in range(n_inst)]
    m.enable_timing(True)
    inst_step = step_b if a.mode == "b" else step_serial
    latency_ms = time_steps(inst_step, n_inst, 2, events=ev)
    kernel_times = m.timings()
    m.enable_timing(False)
    stages = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(4)] for e in ev]).mean(0)

    # ---- end to end through the C-ABI with HOST buffers: scene struct in, the reference's uchar4 frame out.  Mode A: every rank
    #      calls pm_frame_host_async (three-stage pipeline, exchange inside) and copies ITS row band into one host frame -- pinned
    #      memory, shared between the rank processes at N > 1 -- over its own PCIe link; the host waits for every frame, one frame
    #      behind, and rank 0 also waits until every rank has reported its band of that frame. ----
    e2e_steps = max(3, min(a.steps, 30))
    pipelined = a.mode == "a" and a.passes == 1 and (peers or world == 1)
    pending = []
    if pipelined:
        host_frames, done_words, shm_keep = pmdist.shared_host_frames(3, W * H * 4, world)
    else:
        host_frames = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()]
    n_sub = [0]

    def e2e_retire():
        tk, n = pending.pop(0)
        m.frame_wait(tk)                          # this rank's band of frame n is in host memory
        if world > 1:
            done_words[n % 3, rank] = n + 1
            if rank == 0:                         # ... and so is every other rank's
                while int(done_words[n % 3].min()) < n + 1:
                    pass

    def e2e_step():
        m.set_scene(scene)                        # the frame's only host input: the scene / parameter block
        if pipelined:
            n = n_sub[0]; n_sub[0] += 1
            pending.append((m.frame_async(W, H, host_frames[n % 3], t=0.0, emit=True, interp=False, media=True), n))
            if len(pending) > 2:
                e2e_retire()                      # frame f-2 is complete in host memory before frame f+1 is submitted
        else:
            step()
            if after: after()
            if rank == 0:
                fb = keep["fb"][0] if a.mode == "b" else rgba
                host_frames[0].copy_(fb, non_blocking=True)
            torch.cuda.synchronize()

    def e2e_drain():
        while pending:
            e2e_retire()

    for _ in range(3):
        e2e_step()
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_drain()                                   # the last frame is in host memory too
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    clocks = sampler.stop() if sampler else None
    peer_ok = True
    if peers:
        try:
            m.peer_status()
        except pmb200.PmError:
            peer_ok = False

    # ---- extras carried by the same line, so that the driver's scaling run records them at every N: Mode B at N GPUs
    #      (BASELINE config 4, k-NN estimator) and progressive photon mapping (config 5: 64 passes x 16M photons, 3840x2160) ----
    extras = {}
    if a.mode == "a" and a.passes == 1 and not a.no_extras and NP >= (1 << 20):
        try:
            p5 = 64
            W5, H5 = 3840, 2160
            y50, y51 = pmdist.row_band_uneven(H5, rank, world)
            sc5 = pmb200.default_scene(sz_img=H5); sc5.cam_ox = -(W5 - H5) / 2.0
            m.set_scene(sc5); m.set_energy_scale(10000.0 / NP / p5)
            fb5 = torch.zeros((H5, W5, 4), dtype=torch.uint8, device="cuda")

            def step5():
                m.clear_map()
                for _ in range(p5):               # pass p continues the MWC stream: a fresh direction table per pass
                    m.init_random_numbers()
                    m.trace(0.0, media=True)
                if world > 1 and not peers:
                    pmdist.allreduce_accumulators(pmdist.device_tensor(m.accumulators()[0], acc_n, "<i8"))
                m.build_map()
                m.render_device(W5, H5, 0.0, False, True, rgba=fb5, y0=y50, y1=y51)
                if peers:
                    m.peer_barrier()
            ms5 = time_steps(step5, 3, 1)
            extras["config5_progressive"] = {"workload": "64 passes x %d photons, 3840x2160, Mode A, row bands left on their GPUs" % NP,
                                             "ms_per_frame": ms5, "photons_per_s": p5 * NP / (ms5 * 1e-3), "steps": 3}
            m.set_scene(scene); m.set_energy_scale(10000.0 / NP)
            del fb5
        except Exception as ex:
            extras["config5_progressive"] = {"failed": repr(ex)}
        if world > 1:
            try:
                m.set_record_capacity(int(2.6 * (last - first)) + 4096)
                keep["fb"] = (torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda"), torch.zeros((H, W, 4), dtype=torch.float32, device="cuda"))
                evb = [_events(torch, 5) for _ in range(3)]
                msb = time_steps(step_b, 3, 1, events=evb)
                st_b = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(4)] for e in evb]).mean(0)
                extras["mode_b_config4"] = {"workload": "%d photons, k=%d, 11 k-NN gathers per pixel, %dx%d, records all-gathered, trees on every rank, "
                                                        "rows interleaved" % (NP, a.knn, W, H), "ms_per_frame": msb,
                                            "gather_queries_per_s": W * H * 11 / (float(st_b[3]) * 1e-3),
                                            "stages_ms": {"trace_with_records": float(st_b[0]), "allgather_surface_records": float(st_b[1]),
                                                          "allgather_volume_records||build_surface, build_volume": float(st_b[2]),
                                                          "knn_render(+sum)": float(st_b[3])}, "steps": 3}
                keep.clear()
            except Exception as ex:
                extras["mode_b_config4"] = {"failed": repr(ex)}
            torch.cuda.empty_cache()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
        n_local = (last - first) * a.passes
        kern = {k: {"avg_ms": v[0] / v[1], "launches_per_step": v[1] / n_inst} for k, v in kernel_times.items() if v[1]}
        dom = max(kern, key=lambda k: kern[k]["avg_ms"] * kern[k]["launches_per_step"])
        dom_ms = kern[dom]["avg_ms"]
        props = torch.cuda.get_device_properties(local)
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        line = {
            "metric": METRIC if a.mode == "a" else METRIC.replace("(trace", "Mode B k=%d (trace" % a.knn), "value": ms_per_step, "unit": "ms",
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, {"parallelism": "photon-range x%d + row-band x%d" % (world, world),
                                          "exchange": ("peer memory (pm_peer.cu, CUDA IPC)" if peers else "NCCL all-reduce") if world > 1 else "none",
                                          "frames_in_flight": 1 if (a.no_overlap or a.mode == "b") else 3, "trace_sms": trace_sms or props.multi_processor_count}),
            "photons_per_s": NP * a.passes / (ms_per_step * 1e-3), "pixels_per_s": W * H / (ms_per_step * 1e-3),
            "frame_latency_ms": latency_ms,
            "stages_ms": ({"clear+trace": float(stages[0]), "allreduce": (kern.get("peer_reduce_kernel", {}).get("avg_ms", 0.0) if peers else float(stages[1])),
                           "build_map+tables": float(stages[2]) - (kern.get("peer_reduce_kernel", {}).get("avg_ms", 0.0) if peers else 0.0),
                           "render(+assemble)": float(stages[3]),
                           "note": "one frame on one stream (the instrumented pass; event records between the stages add a few us each); "
                                   "allreduce = peer_reduce_kernel incl. waiting for the slowest rank"} if a.mode == "a" else
                          {"trace_with_records": float(stages[0]), "allgather_surface_records": float(stages[1]),
                           "allgather_volume_records||build_surface, build_volume": float(stages[2]),
                           "knn_render(+sum)": float(stages[3])}),
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": C.sizeof(pmb200.Scene) * world,
                    "d2h_bytes_per_step": W * H * 4, "steps": e2e_steps,
                    "what": ("pm_frame_host_async + pm_frame_wait per frame on every rank: scene struct in, emit (+ exchange) + render, each rank's "
                             "row band of the reference's uchar4 frame copied into one pinned host frame over its own PCIe link, up to three "
                             "frames in flight, every frame waited for" if pipelined else
                             "step + D2H of the assembled uchar4 frame on rank 0, synchronous")},
            "kernels": kern,
            "clocks": clocks,
        }
        if peers:
            line["peer_exchange_ok"] = peer_ok
        if dom == "trace_kernel":
            # the Mode A trace keeps no records, so it is bound by FP32 issue (~400 flop per photon for the
            # intersection / reflection / voxel arithmetic), not by HBM (12 B per photon)
            flops = 400.0 * n_local
            peak_tf = props.multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
            ach_tf = flops / (dom_ms * 1e-3) / 1e12
            prof = {}
            try:
                prof = json.load(open(os.path.join(ROOT, "profiles", "r2_trace_ncu.json")))
            except Exception:
                pass
            same = prof.get("photons") == NP and world == 1
            line["roofline"] = {
                "kernel": dom, "bound": "fp32_issue", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                "traffic": prof.get("dram_bytes_per_launch") if same else None,
                "algorithmic_flop_per_photon": 400.0, "photons_per_launch": n_local, "avg_launch_ms": dom_ms,
                "peak_source": "%d SMs x 128 FP32 lanes x 2 flop x %.0f MHz (sampled)" % (props.multi_processor_count, sm_mhz),
                "hbm": {"achieved": 12.0 * n_local / (dom_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": 12.0 * n_local / (dom_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src,
                        "algorithmic_bytes_per_photon": 12.0},
                "ncu": ({k: prof.get(k) for k in ("issue_slots_busy_pct", "warp_instructions", "thread_instructions_per_photon",
                                                  "active_lanes_per_instruction", "source")} if same else None),
                "note": "FP32-issue roofline of SURVEY.md 8(d) (400 flop per photon); the HBM figure (one 12 B direction per photon) is "
                        "given beside it -- the kernel is not HBM-bound; traffic / ncu.* are read from the committed ncu capture of this "
                        "configuration (profiles/), not measured live"}
        else:
            alg = {"render_kernel": 20.0 * W * rows, "knn_render_kernel": 1632.0 * W * rows * 11}
            bytes_ = alg.get(dom, 0.0)
            line["roofline"] = {"kernel": dom, "bound": "hbm", "achieved": bytes_ / (dom_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                "frac": bytes_ / (dom_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "peak_source": peak_src}
        line.update(extras)
        if not a.no_cpu_baseline and world == 1 and a.mode == "a":
            try:
                from oracle import refhost
                v, d = cpu_reference_frame_ms(a, *((1, 1) if refhost.available() else (16, 8)))
                line["cpu_baseline"] = {"value": v, "unit": "ms", "cores": d["cores"], "kind": d["kind"], "sample": d["sample"],
                                        "emit_ms": d["emit_ms"], "render_ms": d["render_ms"]}
            except Exception as ex:   # the baseline is reported, never required for the measurement itself
                line["cpu_baseline"] = {"value": None, "unit": "ms", "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}
        if not a.no_mode_b and not a.no_extras and world == 1 and a.mode == "a":
            try:
                line["mode_b"] = mode_b_numbers(pmb200, torch, a, local)
            except Exception as ex:
                line["mode_b"] = {"failed": repr(ex)}
        if not a.no_ref_cuda and not a.no_extras and world == 1 and a.mode == "a" and a.passes == 1:
            try:
                line["config2"] = config2_numbers(pmb200, torch, local)
            except Exception as ex:
                line["config2"] = {"failed": repr(ex)}
        if not a.no_ref_cuda and world == 1 and a.mode == "a":
            try:
                fb = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
                rc = time_reference_cuda(a, m.get_random_table(), fb)
                line["reference_cuda_kernel"] = rc
                if "ms_per_frame" in rc:
                    line["vs_reference_cuda_kernel"] = {"speedup": rc["ms_per_frame"] / ms_per_step, "e2e_speedup": rc["ms_per_frame"] / e2e_ms,
                                                        "note": "the reference's own CUDA kernels (sm_100a build) on this GPU, same configuration, timed in this run"}
            except Exception as ex:
                line["reference_cuda_kernel"] = {"unavailable": repr(ex)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
    if pipelined and shm_keep is not None:
        del host_frames, done_words
        shm_keep.close()
    if peers:
        m.shared_close(rgba_ptr, rgba_opened)
        m.peer_disconnect()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

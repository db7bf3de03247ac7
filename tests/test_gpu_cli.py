"""Process-level behaviour: the legacy symbols' print-and-exit error path (checkCUDAError, PMK:49-55) and the headless driver
(cuda-photon-mapper_b200/pm_headless, the replacement for simpleGLMain.cpp / callbacksPBO.cpp) with its epsilon / threshold
image compare (the SDK sample's regression mode, simpleGL.cpp:354-368).  Each case runs in a subprocess."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADLESS = os.path.join(ROOT, "cuda-photon-mapper_b200", "pm_headless")


def _py(code, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r)\n%s" % (ROOT, code)], env=e, capture_output=True,
                          text=True, timeout=300)


def test_legacy_launcher_prints_and_exits_on_a_cuda_error():
    """launch_photon_mapping_kernel with a pointer that is not device memory: the kernel faults, the launcher's
    cudaThreadSynchronize + checkCUDAError equivalent prints `Cuda error: <msg>: <str>.` and exits with EXIT_FAILURE -- the
    reference's contract (PMK:49-55, :1563-1565), not a return code."""
    r = _py("import pmb200\n"
            "pmb200.launch_init_random_numbers_kernel()\n"
            "pmb200.launch_emit_photons_kernel(None, 64, 64, 0.0, False, False)\n"
            "pmb200.launch_photon_mapping_kernel(0x10, 64, 64, 0.0, False, False)\n"
            "print('not reached')\n")
    assert r.returncode == 1, (r.returncode, r.stderr)
    assert "not reached" not in r.stdout
    assert r.stderr.strip().splitlines()[-1].startswith("Cuda error: photon_mapping_kernel failed!: "), r.stderr
    assert r.stderr.strip().endswith(".")


def test_default_context_failure_prints_and_exits():
    """The same convention when the default context cannot be set up (photon count below the three table rows the medium walk reads)."""
    r = _py("import pmb200\npmb200.launch_init_random_numbers_kernel()\nprint('not reached')\n", env={"PMB200_NR_PHOTONS": "2"})
    assert r.returncode == 1 and "not reached" not in r.stdout
    assert r.stderr.strip().startswith("Cuda error: PMB200_NR_PHOTONS: "), r.stderr


def test_legacy_launchers_succeed_in_display_order():
    r = _py("import torch, pmb200\n"
            "fb = torch.zeros((64, 64, 4), dtype=torch.uint8, device='cuda')\n"
            "pmb200.launch_init_random_numbers_kernel()\n"
            "pmb200.launch_emit_photons_kernel(fb, 64, 64, 0.0, False, True)\n"
            "pmb200.launch_photon_mapping_kernel(fb, 64, 64, 0.0, False, True)\n"
            "print('sum', int(fb.sum()))\n")
    assert r.returncode == 0, r.stderr
    assert int(r.stdout.split()[-1]) > 0


def _read_ppm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P6"
        w, h = (int(x) for x in f.readline().split())
        assert f.readline().strip() == b"255"
        return np.frombuffer(f.read(), np.uint8).reshape(h, w, 3).copy()


def test_headless_driver_writes_frames_and_compares(tmp_path, pm):
    """pm_headless: two animated frames as PPM + PFM; the PPM equals the uchar4 frame of pm_frame_host for the same parameters;
    --compare passes against its own output, fails (exit status 1) against a visibly different image, and honours eps / threshold."""
    if not os.path.exists(HEADLESS):
        pytest.skip("pm_headless not built")
    out = str(tmp_path / "f")
    args = [HEADLESS, "--photons", "20000", "--width", "160", "--height", "120", "--media", "1", "--time", "0.5", "--dt", "0.25"]
    r = subprocess.run(args + ["--frames", "2", "--out", out, "--pfm"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    img0, img1 = _read_ppm(out + "_0000.ppm"), _read_ppm(out + "_0001.ppm")
    assert img0.shape == (120, 160, 3) and img0.any() and (img0 != img1).any()
    assert os.path.getsize(out + "_0000.pfm") > 160 * 120 * 12
    # the same frame through the C-ABI from Python
    m = pm.PhotonMapper(n_photons=20000)
    sc = pm.default_scene(sz_img=120); sc.cam_ox = -(160 - 120) / 2.0
    m.set_scene(sc); m.set_energy_scale(10000.0 / 20000); m.init_random_numbers()
    u8 = np.zeros((120, 160, 4), np.uint8)
    m.frame(160, 120, 0.5, True, False, True, out_u8=u8)
    assert np.array_equal(u8[..., :3], img0)
    m.close()
    one = args + ["--frames", "1", "--out", str(tmp_path / "g")]
    r = subprocess.run(one + ["--compare", out + "_0000.ppm"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASS" in r.stdout, (r.stdout, r.stderr)
    bad = str(tmp_path / "bad.ppm")
    with open(bad, "wb") as f:
        f.write(b"P6\n160 120\n255\n" + (255 - img0).tobytes())
    r = subprocess.run(one + ["--compare", bad], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "FAIL" in r.stdout
    r = subprocess.run(one + ["--compare", bad, "--eps", "255"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASS" in r.stdout          # nothing differs by more than 255
    r = subprocess.run(one + ["--compare", bad, "--threshold", "1.0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0                                  # any fraction of differing values is tolerated

"""CPU: host-side pieces of bench.py that must not break the one JSON line (no GPU needed)."""
import importlib.util
import json
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class _FakeProc:
    def terminate(self):
        pass


def test_clock_sampler_window_and_reasons():
    b = _bench()
    s = b.ClockSampler(0)
    s.proc = _FakeProc()
    t0 = time.perf_counter()
    row = lambda sm, pw, cap: ["0", str(sm), "1965", str(pw), "0x4", "Not Active", "Not Active", "Not Active", cap]
    s.rows = [(t0 - 1.0, row(1965, 250.0, "Not Active")),           # before the window: ignored
              (t0 + 0.01, row(1700, 990.0, "Active")), (t0 + 0.03, row(1680, 1001.5, "Active")),
              (t0 + 5.0, row(1965, 200.0, "Not Active"))]           # after: ignored
    c = s.stop(t0, t0 + 0.04)
    assert c["samples"] == 2 and c["sm_mhz"] == 1690.0 and c["sm_max_mhz"] == 1965.0
    assert c["reasons"] == ["sw_power_cap"] and c["power_w_max"] == 1001.5 and c["window"] == "timed region"
    # a region shorter than the sampling period falls back to the nearest samples and says so
    s2 = b.ClockSampler(0)
    s2.proc = _FakeProc()
    s2.rows = [(t0 - 0.05, row(1965, 300.0, "Not Active"))]
    c2 = s2.stop(t0, t0 + 0.001)
    assert c2["samples"] == 1 and c2["window"].startswith("nearest")
    # no nvidia-smi at all
    s3 = b.ClockSampler(0)
    assert s3.stop(t0, t0 + 1)["sm_mhz"] is None


def test_peaks_and_constants():
    b = _bench()
    p = b.peaks()
    assert p["hbm_gbs"] > 1000 and p["tflops_sustained"] <= p["tflops_burst"]
    # algorithmic work per stereo pair: decoder 7595.9 MF, fused soft-argmax + DLT 623 220 B
    dec = sum(b.FLOP_PER_PAIR[k] for k in ("deconv1", "deconv2", "deconv3", "final_1x1"))
    assert abs(dec - 7595.9e6) < 1e5 and b.SOFTARGMAX_DLT_BYTES_PER_POSE == 2 * 19 * 4096 * 4 + 96 + 304 + 228
    t = b.ncu_traffic()
    assert "fp32" in t and t["fp32"]["deconv3"] > 5e8 and json.dumps(t)

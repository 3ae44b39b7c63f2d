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
    s.mode, s.proc = "nvidia-smi", _FakeProc()
    t0 = time.perf_counter()
    row = lambda t, sm, pw, cap: (t, float(sm), 1965.0, pw, ["sw_power_cap"] if cap else [])
    s.rows = [row(t0 - 1.0, 1965, 250.0, False),                    # before the window: ignored
              row(t0 + 0.01, 1700, 990.0, True), row(t0 + 0.03, 1680, 1001.5, True),
              row(t0 + 5.0, 1965, 200.0, False)]                    # after: ignored
    c = s.stop(t0, t0 + 0.04)
    assert c["samples"] == 2 and c["sm_mhz"] == 1690.0 and c["sm_max_mhz"] == 1965.0 and c["sm_mhz_min"] == 1680.0
    assert c["reasons"] == ["sw_power_cap"] and c["power_w_max"] == 1001.5 and c["window"] == "timed region"
    # a region shorter than the sampling period falls back to the nearest samples and says so
    s2 = b.ClockSampler(0)
    s2.mode, s2.proc = "nvidia-smi", _FakeProc()
    s2.rows = [row(t0 - 0.05, 1965, 300.0, False)]
    c2 = s2.stop(t0, t0 + 0.001)
    assert c2["samples"] == 1 and c2["window"].startswith("nearest")
    # the nvidia-smi fallback parses its csv rows into the same tuples
    s4 = b.ClockSampler(0)

    class _P:
        stdout = ["0, 1700, 1965, 990.0, 0x4, Not Active, Not Active, Not Active, Active\n", "garbage\n"]
    s4.proc = _P()
    s4._read()
    assert len(s4.rows) == 1 and s4.rows[0][1:] == (1700.0, 1965.0, 990.0, ["sw_power_cap"])
    # no NVML / nvidia-smi at all
    s3 = b.ClockSampler(0)
    assert s3.stop(t0, t0 + 1)["sm_mhz"] is None


def test_both_arms_share_one_config():
    """The driver compares the `config` of `--impl ours` and `--impl reference` (same_config): one function."""
    b = _bench()
    c = b.bench_config(64, 1)
    assert c == b.bench_config(64, 1) and c["workload"].startswith("CDRNet head") and "BASELINE configs[1]" in c["workload"]
    assert b.bench_config(64, 8)["parallelism"] == "dp8" and json.dumps(c)


def test_peaks_and_constants():
    b = _bench()
    p = b.peaks()
    assert p["hbm_gbs"] > 1000 and p["tflops_sustained"] <= p["tflops_burst"]
    # algorithmic work per stereo pair: decoder 7595.9 MF, fused soft-argmax + DLT 623 220 B
    dec = sum(b.FLOP_PER_PAIR[k] for k in ("deconv1", "deconv2", "deconv3", "final_1x1"))
    assert b.FLOP_PER_PAIR["deconv3_tail"] == b.FLOP_PER_PAIR["deconv3"] + b.FLOP_PER_PAIR["final_1x1"]
    assert abs(dec - 7595.9e6) < 1e5 and b.SOFTARGMAX_DLT_BYTES_PER_POSE == 2 * 19 * 4096 * 4 + 96 + 304 + 228
    t = b.ncu_traffic()
    top = t["fp32"].get("deconv3_tail", t["fp32"].get("deconv3"))          # r02: fused tail; r01: deconv3
    assert "fp32" in t and top > 1e8 and json.dumps(t)

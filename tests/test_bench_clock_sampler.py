"""bench.py's clock sampler without a GPU: a stand-in `pynvml` on PYTHONPATH lets the helper process run.  Checked: samples are
taken only while a timed region holds the gate (the helper's NVML queries must not run during the e2e leg: they contend with the
driver's allocation calls), at the requested period, and the summary names the throttle reasons from the NVML bit mask."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

FAKE = '''
NVML_CLOCK_SM = 1
def nvmlInit(): pass
def nvmlDeviceGetHandleByUUID(u): raise RuntimeError("no such device")
def nvmlDeviceGetHandleByIndex(i): return i
def nvmlDeviceGetMaxClockInfo(h, k): return 1965
def nvmlDeviceGetClockInfo(h, k): return 1800
def nvmlDeviceGetPowerUsage(h): return 350000
def nvmlDeviceGetCurrentClocksEventReasons(h): return 0x4 | 0x1      # sw_power_cap | gpu_idle
'''


def test_sampler_polls_only_inside_timed_regions(tmp_path, monkeypatch):
    (tmp_path / "pynvml.py").write_text(FAKE)
    monkeypatch.setenv("PYTHONPATH", str(tmp_path))
    monkeypatch.syspath_prepend(str(tmp_path))
    sys.modules.pop("pynvml", None)
    import bench
    monkeypatch.setattr(bench.ClockSampler, "_helpers", {})
    s = bench.ClockSampler(3)
    proc, path = bench.ClockSampler._helpers[3]
    try:
        time.sleep(0.6)                                    # helper start-up; no gate: nothing may be sampled
        assert proc.poll() is None
        assert Path(path).read_text() == ""
        with s:
            time.sleep(0.15)
        out = s.summary()
        assert out["samples"] >= 5 and out["sm_mhz"] == 1800.0 and out["sm_max_mhz"] == 1965.0
        assert out["reasons"] == ["sw_power_cap"] and out["power_w_max"] == 350.0
        n_after = len(Path(path).read_text().splitlines())
        time.sleep(0.1)                                    # gate released: the file stops growing
        assert len(Path(path).read_text().splitlines()) <= n_after + 1
        slow = bench.ClockSampler(3, period_ms=50.0)       # same helper, longer period (the reference arm)
        with slow:
            time.sleep(0.3)
        assert 2 <= slow.summary()["samples"] <= 8
    finally:
        proc.terminate()
        sys.modules.pop("pynvml", None)

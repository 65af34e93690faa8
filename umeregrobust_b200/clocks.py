"""SM clock / throttle-reason sampling during a timed region (bench.py's `clocks` record).

NVML (nvidia_ml_py) is polled from a thread every few milliseconds, so that even a sub-second timed
region gets tens of samples; `nvidia-smi -lms` is the fallback when NVML cannot be loaded."""
import subprocess
import threading
import time

_REASONS = {
    "hw_slowdown": 0x0000000000000008,
    "sw_power_cap": 0x0000000000000004,
    "hw_thermal_slowdown": 0x0000000000000040,
    "hw_power_brake_slowdown": 0x0000000000000080,
    "sw_thermal_slowdown": 0x0000000000000020,
}


class ClockSampler:
    def __init__(self, gpu_index, period_s=0.005, uuid=None):
        self.gpu = int(gpu_index)
        self.uuid = uuid            # CUDA device UUID: NVML enumerates physical GPUs, CUDA the visible ones
        self.period = period_s
        self.samples = []          # (t, sm_mhz, reasons_bitmask)
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        self._smi = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = None
            if self.uuid:
                for cand in ("GPU-" + str(self.uuid), str(self.uuid)):
                    try:
                        self._h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        try:
                            self._h = pynvml.nvmlDeviceGetHandleByUUID(cand)
                            break
                        except Exception:
                            self._h = None
            if self._h is None:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None
            try:
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
                self._smi = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                              "--format=csv,noheader,nounits", "-lms", "50"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self._thread = threading.Thread(target=self._poll_smi, daemon=True)
                self._thread.start()
            except Exception:
                self._smi = None

    def _poll_nvml(self):
        n = self._nvml
        while not self._stop.is_set():
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                self.samples.append((time.time(), mhz, mask))
            except Exception:
                pass
            time.sleep(self.period)

    def _poll_smi(self):
        for line in self._smi.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.sm_max = float(f[1])
                self.samples.append((time.time(), float(f[0]), int(f[2], 16) if f[2].startswith("0x") else 0))
            except Exception:
                pass

    def stop(self, t0, t1):
        self._stop.set()
        if self._smi is not None:
            time.sleep(0.1)
            self._smi.terminate()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "samples": 0, "reasons": ["no clock samples"]}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        mhz = sorted(s[1] for s in inside)
        mask = 0
        for s in inside:
            mask |= s[2]
        reasons = sorted(k for k, bit in _REASONS.items() if mask & bit)
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.sm_max, "samples": len(inside), "reasons": reasons}

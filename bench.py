#!/usr/bin/env python
"""bench.py -- 4Kp60 8-bit preset-M10 hot-path throughput (CTUs/s, fps) on N B200s, next to the
reference's CPU path on the same box.

A "step" = one 3840x2160 frame through the hot path (stages listed in config.stages) with the
per-picture parameters preset M10 derives for that frame's temporal layer; consecutive steps walk the
8-frame hierarchical mini-GOP [0,3,2,3,1,3,2,3] (3 hierarchical levels, EbApi.h:222-226).
  value : frames resident in HBM (16-frame ring > L2), device-timed with CUDA events, max over ranks
  e2e   : the same step through the C ABI with HOST (pinned) buffers: H2D of the frame, kernels, D2H of results
N > 1   : independent closed GOPs / streams per GPU -- weak scaling, no data-path collective.
--impl reference : the reference's own CPU implementation of the same stages (oracle/_ref, all host threads).
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
for p in (os.path.join(ROOT, "svt-hevc_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

W, H = 3840, 2160
N_LCU = 60 * 34
TL_PATTERN = [0, 3, 2, 3, 1, 3, 2, 3]
RING = 16
# algorithmic HBM bytes per CTU per reference list for ME: src Y 4096 + src 1/16 256 + ref Y 4096 + ref 1/16 256 + results 680
ME_BYTES_PER_CTU_PER_LIST = 9384


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_frames(n, seed):
    import common
    v = common.SynthVideo(W, H, seed=seed)
    return [v.luma(i) for i in range(n)]


def cfg_for(step, S):
    tl = TL_PATTERN[step % 8]
    return S.me_config(W, H, enc_mode=10, temporal_layer=tl, num_lists=2, ref_poc=(0, 16)), tl


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import svtb200 as S

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S.check(S.lib().svtb200_init(C.c_int(local)))
    stream = torch.cuda.current_stream().cuda_stream

    lumas = make_frames(RING, seed=0x5F7C0001 + rank)
    pinned = [torch.from_numpy(l).pin_memory() for l in lumas]
    ring = [S.DevPaPicture(W, H) for _ in range(RING)]
    for d, p in zip(ring, pinned):
        d.upload(p.data_ptr(), stream)
    res_bytes = N_LCU * 85 * 24
    d_res = torch.empty(res_bytes, dtype=torch.uint8, device="cuda")
    h_res = torch.empty(res_bytes, dtype=torch.uint8).pin_memory()
    e2e_pic = S.DevPaPicture(W, H)
    torch.cuda.synchronize()

    def refs_for(i):   # mini-GOP style references: nearest earlier / later ring entries by temporal distance
        tl = TL_PATTERN[i % 8]
        d = {0: 8, 1: 4, 2: 2, 3: 1}[tl]
        return ring[(i - d) % RING], ring[(i + d) % RING]

    def step_resident(i, ev=None):
        cfg, _ = cfg_for(i, S)
        r0, r1 = refs_for(i)
        if ev: ev[0].record()
        S.me_picture_dev(cfg, ring[i % RING], r0, r1, d_res.data_ptr(), stream)
        if ev: ev[1].record()

    def step_e2e(i):
        cfg, _ = cfg_for(i, S)
        r0, r1 = refs_for(i)
        e2e_pic.upload(pinned[i % RING].data_ptr(), stream)                      # H2D luma + device pad/decimate
        S.me_picture_dev(cfg, e2e_pic, r0, r1, d_res.data_ptr(), stream)
        h_res.copy_(d_res, non_blocking=True)                                    # D2H results
        torch.cuda.current_stream().synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step_resident(i)
    launches0 = S.lib().svtb200_launch_count()
    sampler = ClockSampler(local); sampler.start()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for i in range(args.steps):
        step_resident(args.warmup + i, kev[i])
    t1.record()
    barrier()
    sampler.stop_flag = True
    launches = S.lib().svtb200_launch_count() - launches0
    ms = t0.elapsed_time(t1)
    kms = [a.elapsed_time(b) for a, b in kev]

    # e2e: host buffers in, results out, copies inside the timed region
    for i in range(args.warmup):
        step_e2e(i)
    barrier()
    w0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(args.warmup + i)
    barrier()
    e2e_s = time.perf_counter() - w0

    tms = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tms[0]), float(tms[1])
    if rank == 0:
        peak, peak_src = peaks()
        ctus = N_LCU * args.steps * world
        k_avg_s = float(np.mean(kms)) * 1e-3
        alg_bytes = ME_BYTES_PER_CTU_PER_LIST * 2 * N_LCU
        achieved = alg_bytes / k_avg_s / 1e9
        out = {
            "metric": "4Kp60 8-bit preset-M10 hot-path CTUs/sec (fps = value / 2040)", "value": ctus / (ms * 1e-3), "unit": "CTU/s",
            "fps": ctus / (ms * 1e-3) / N_LCU,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: 3840x2160 8-bit 4:2:0, preset M10 parameters, mini-GOP temporal layers [0,3,2,3,1,3,2,3], 2 reference lists",
                       "stages": ["me_picture(HME-L0 one quadrant + 85-PU full-pel + integer bipred)"],
                       "l2": "16-frame ring of resident 4K pictures (~175 MB of luma planes) > 126 MB L2", "parallelism": f"stream-per-GPU x{world}"},
            "e2e": {"value": ctus / (e2e_ms * 1e-3), "unit": "CTU/s", "fps": ctus / (e2e_ms * 1e-3) / N_LCU,
                    "h2d_bytes_per_step": W * H, "d2h_bytes_per_step": res_bytes},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"kernel": "me_picture_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": k_avg_s * 1e3},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_reference(steps=8, max_threads=None)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own MotionEstimateLcu + SIMD kernels, all host threads
# ------------------------------------------------------------------------------------------------
def cpu_reference(steps, max_threads):
    import common
    import svtb200 as S
    from concurrent.futures import ThreadPoolExecutor
    ref = common.load_ref()
    kind = "reference"
    fn_name = "svtref_me_picture_rows"
    lib = ref
    if ref is None:
        raise RuntimeError("oracle/_ref missing: build it with oracle/build_ref.sh where /root/reference exists")
    nthreads = max_threads or (os.cpu_count() or 1)
    lumas = make_frames(3, seed=0x5F7C0001)
    pics = [S.HostPaPicture(l) for l in lumas]
    fn = getattr(lib, fn_name)
    fn.restype = C.c_int
    res = np.zeros((N_LCU, 85), dtype=S.ME_PU_RESULT_DTYPE)
    rows = 34
    chunks = [(r, r + 1) for r in range(rows)]

    def work(args):
        cfg, r0, r1 = args
        return fn(C.byref(cfg), C.byref(pics[1].struct), C.byref(pics[0].struct), C.byref(pics[2].struct), res.ctypes.data_as(C.c_void_p),
                  None, None, None, C.c_int(1), C.c_int(r0), C.c_int(r1))

    with ThreadPoolExecutor(nthreads) as ex:
        cfg0, _ = cfg_for(0, S)
        assert all(r == 0 for r in ex.map(work, [(cfg0, a, b) for a, b in chunks]))          # warm-up frame
        cfgs = [cfg_for(s, S)[0] for s in range(steps)]
        t0 = time.perf_counter()
        rc = list(ex.map(work, [(c, a, b) for c in cfgs for a, b in chunks]))                  # frames pipelined like the reference does
        t_total = time.perf_counter() - t0
        assert all(r == 0 for r in rc)
    v = N_LCU * steps / t_total
    return {"value": v, "unit": "CTU/s", "fps": v / N_LCU, "cores": min(nthreads, rows * steps), "kind": kind,
            "sample": f"{steps} full 4K frames (after 1 warm-up) of the same stages: reference MotionEstimateLcu with its AVX2/AVX-512 kernels, one LCU row per task"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 16))
    t0 = time.perf_counter()
    cb = cpu_reference(steps=steps, max_threads=None)
    out = {"impl": "reference", "metric": "4Kp60 8-bit preset-M10 hot-path CTUs/sec (fps = value / 2040)", "value": cb["value"], "unit": "CTU/s",
           "fps": cb["fps"], "n_gpus": world, "steps": steps, "warmup": 1, "ms_per_step": 1e3 * N_LCU / cb["value"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": {"workload": "configs[1]: 3840x2160 8-bit 4:2:0, preset M10 parameters, mini-GOP temporal layers [0,3,2,3,1,3,2,3], 2 reference lists",
                      "stages": ["me_picture(HME-L0 one quadrant + 85-PU full-pel + integer bipred)"]},
           "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "CTU/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": time.perf_counter() - t0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

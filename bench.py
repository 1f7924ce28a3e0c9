#!/usr/bin/env python
"""bench.py -- 4Kp60 8-bit preset-M10 hot-path throughput (CTUs/s, fps) on N B200s, next to the reference's CPU path
on the same box.

A "step" = one 3840x2160 4:2:0 frame through the hot path with the per-picture tool schedule preset M10 derives for that
frame's temporal layer; consecutive steps walk the 8-frame hierarchical mini-GOP [0,3,2,3,1,3,2,3]:
    every frame : ME (HME-L0 one quadrant + 85-PU full-pel + bipred, 2 lists) -> open-loop intra search (DC only at M10) -> MC of the chosen PUs
                  -> residual/T/Q/iQ/iT/recon -> SAO statistics
    layer 0 only: + selective half/quarter-pel ME refinement (useSubpelFlag, fractionalSearchModel 1) + deblocking + SAO application
                  (M10 switches sub-pel, DLF and SAO-apply off on layers 1-3: EbPictureDecisionProcess.c:353-361, EbEncDecProcess.c:2036-2054)
Mode decision itself (host, out of scope) is replaced by a fixed stand-in: 32x32 PUs / TUs taking each PU's best ME
candidate; SAO parameters and boundary strengths are fixed per-LCU tables.  Both arms run exactly the same jobs.
  value : frames resident in HBM (16-frame ring > L2), device-timed with CUDA events, max over ranks; consecutive (independent) frames
          alternate between --streams compute streams, per-stage events are taken inside that timed region (`stages`), and once more
          with a single stream (`stages_serial`: what each kernel costs alone)
  e2e   : the same step through the C ABI from HOST (pinned) buffers: H2D of Y/Cb/Cr, kernels, D2H of ME results, OIS candidates,
          quantised coefficients, nz counts and SAO statistics; --inflight frames in flight (copy-in / compute / copy-out streams); coefficients leave as the packed stream of coded TUs
  N > 1 : independent closed GOPs / streams per GPU -- weak scaling, no data-path collective
  --impl reference : the reference's own CPU implementation of the same stages (oracle/_ref, one frame per host thread)
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
COLS, ROWS = 60, 34
N_LCU = COLS * ROWS
BITS, ENC_MODE, FPS, CFG_ID = 8, 10, 60, 2
TL_PATTERN = [0, 3, 2, 3, 1, 3, 2, 3]
RING = 16
PAD = 80
# BASELINE.json configs (index in `configs`, 0-based = id - 1).  2 is the headline the driver runs; 3 / 4 / 5: --config N
CONFIGS = {
    2: dict(w=3840, h=2160, bits=8, enc_mode=10, fps=60, streams=3, gop_exchange=False, graphs=False,
            metric="4Kp60 8-bit preset-M10 hot-path CTUs/sec (fps = value / 2040)",
            workload="configs[1]: 3840x2160 8-bit 4:2:0, preset M10 tool schedule, mini-GOP temporal layers [0,3,2,3,1,3,2,3], 2 reference lists, "
                     "fixed 32x32 PU/TU mode-decision stand-in"),
    3: dict(w=3840, h=2160, bits=10, enc_mode=7, fps=60, streams=3, gop_exchange=False, graphs=True,
            metric="4Kp60 10-bit preset-M7 hot-path CTUs/sec (fps = value / 2040)",
            workload="configs[2]: 3840x2160 10-bit 4:2:0, preset M7 tool schedule (HME level 0 in 2x2 regions + level 1, 16x9 full-pel search, selective sub-pel "
                     "refinement on every layer, full open-loop intra search), 8-bit + 2-bit input planes packed to 16-bit samples on the device, 16-bit MC / "
                     "T-Q / DLF / SAO, mini-GOP temporal layers [0,3,2,3,1,3,2,3], 2 reference lists, fixed 32x32 PU/TU mode-decision stand-in"),
    4: dict(w=7680, h=4320, bits=10, enc_mode=10, fps=50, streams=3, gop_exchange=True, graphs=False,
            metric="8Kp50 10-bit preset-M10 hot-path CTUs/sec, closed GOPs round-robin over the GPUs (fps = value / 8160)",
            workload="configs[3]: 7680x4320 10-bit 4:2:0, preset M10 tool schedule, closed GOPs of 16 pictures, GOP g on rank g % N, per-picture outputs "
                     "(packed coefficients + offsets + nz + ME / OIS / SAO records) sent to rank 0 over NCCL at every GOP boundary inside the timed region"),
    5: dict(w=1920, h=1080, bits=8, enc_mode=10, fps=60, streams=4, gop_exchange=False, graphs=True,
            metric="1080p60 8-bit preset-M10 stream-density hot-path CTUs/sec, 4 concurrent streams per GPU (fps per stream = value / 510 / (4 * N))",
            workload="configs[4]: 4 independent 1920x1080 8-bit streams per GPU (32 on 8 GPUs), preset M10 (the reference rejects M12 at 1080p, "
                     "EbEncHandle.c:2304-2309), one compute stream per video stream, no communication"),
}
METRIC = CONFIGS[2]["metric"]
WORKLOAD = CONFIGS[2]["workload"]
STAGES = ["me_picture (+ sub-pel on layer 0)", "ois_picture", "mc_batch", "tq_batch", "sao_stats_picture (H3 state on layer 0: deblocked + unfiltered strips)",
          "sao_decision_picture (SaoGenerationDecision; its offsets feed sao_apply on layer 0)", "dlf_picture (layer 0)", "sao_apply_picture (layer 0)"]
# algorithmic HBM bytes per CTU of 8-bit samples; sample planes count twice at 10 bits (alg_per_ctu)
ALG = {"me": 9384 * 2, "ois": 4096 + 21 * 5, "mc_uni": 12288, "mc_bi": 18432, "tq": 30720, "dlf": 12864, "sao_stats": 12288 + 3 * 312, "sao_dec": 3 * 312 + 40 + 20, "sao_apply": 12288}


def set_config(cid):
    global W, H, COLS, ROWS, N_LCU, BITS, ENC_MODE, FPS, CFG_ID, METRIC, WORKLOAD
    c = CONFIGS[cid]
    W, H, BITS, ENC_MODE, FPS, CFG_ID = c["w"], c["h"], c["bits"], c["enc_mode"], c["fps"], cid
    COLS, ROWS = (W + 63) // 64, (H + 63) // 64
    N_LCU = COLS * ROWS
    METRIC, WORKLOAD = c["metric"], c["workload"]
    return c


# SAO decision inputs: CABAC bit estimates (1 bit = 32768) of the magnitudes MdRateEstimationContext_t holds for saoMergeFlagBits / saoTypeIndexBits /
# saoOffsetTrunUnaryBits, and fullLambda / fullChromaLambdaSao of a QP-32 picture (lambdaModeDecision* tables, EbLambdaRateTables.h): the same values on both arms
SAO_LAMBDA = (439162, 350000)


def sao_rates():
    import svtb200 as S
    t = S.SaoRateTables()
    t.merge_flag_bits[:] = [9000, 70000]
    t.type_index_bits[:] = [14000, 80000, 80000, 80000, 80000, 81000]
    t.offset_trun_unary_bits[:] = [32768 * (k + 1) if k < 7 else 32768 * 7 for k in range(8)]
    return t


def alg_per_ctu():
    """algorithmic bytes per CTU of the current configuration: ME / OIS read the 8-bit planes at any bit depth, the EncDec stages move 16-bit samples at 10 bits"""
    k = 2 if BITS == 10 else 1
    a = dict(ALG)
    if ENC_MODE == 7:
        a["me"] = (9384 + 2048) * 2            # + the quarter-resolution planes of HME level 1
    for n in ("mc_uni", "mc_bi", "dlf", "sao_apply"):
        a[n] = ALG[n] * k
    a["tq"] = (6144 + 6144) * k + 12288 + 6144 * k
    a["sao_stats"] = 12288 * k + 3 * 312
    a["pack"] = 6144 + 6144 + 12288
    return a


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (threads + first-touch of everything it pins afterwards) to the CPUs of the NUMA node its GPU hangs off (NVML's ideal CPU affinity),
    intersected with the CPUs the process may use.  Returns a record for the JSON line."""
    rec = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = sorted(cpus & allowed)
        if use:
            os.sched_setaffinity(0, use)
            rec = {"bound": True, "cpus": f"{use[0]}-{use[-1]} ({len(use)})", "allowed_before": len(allowed)}
            try:
                nodes = [n for n in os.listdir("/sys/devices/system/node") if n.startswith("node")]
                for n in nodes:
                    with open(f"/sys/devices/system/node/{n}/cpulist") as f:
                        lst = f.read().strip()
                    first = int(lst.split(",")[0].split("-")[0])
                    if first in use or use[0] in _cpulist(lst):
                        rec["numa_node"] = n
                        break
            except Exception:
                pass
    except Exception as e:      # noqa: BLE001
        rec["error"] = f"{type(e).__name__}: {e}"
    return rec


def _cpulist(s):
    out = set()
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-"); out.update(range(int(a), int(b) + 1))
        elif part:
            out.add(int(part))
    return out


def kernel_sources_sha16():
    """sha256 (first 16 hex digits) of the CUDA sources: stamps the ncu traffic file with the code it was captured from"""
    import hashlib
    m = hashlib.sha256()
    d = os.path.join(ROOT, "svt-hevc_b200", "csrc")
    for n in sorted(os.listdir(d)):
        with open(os.path.join(d, n), "rb") as f:
            m.update(n.encode()); m.update(f.read())
    return m.hexdigest()[:16]


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of each stage's kernel from the committed `ncu --set full` capture of this command
    (profiles/r02_traffic.json, written by tools/ncu_traffic.py on the GPU box).  The file is stamped with the hash of the kernel sources it was
    captured from: a stale file yields null, never a number of other code."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        if t.get("sources_sha16") == kernel_sources_sha16():
            return t["bytes_per_launch"], "profiles/r02_traffic.json"
        return {}, "profiles/r02_traffic.json is stale (captured from other kernel sources)"
    except Exception:
        return {}, None


# VABSDIFF4 issue rate measured on B200 by tools/ubench_pipes.cu (profiles/r02_ubench_pipes.txt): 0.49 warp-instructions / clock / SM sub-partition,
# shared with SHF / PRMT / IADD3 / LOP3 / IMNMX (one half-rate ALU pipe)
VABSDIFF4_PER_CLK_PER_SMSP = 0.49
HME_MULT = {0: 200, 1: 140, 2: 100, 3: 70}          # HME_LEVEL_0_SEARCH_AREA_MULTIPLIER_X/_Y, 4 hierarchical levels (EbDefinitions.h:1322-1331)


def me_sad_warp_instr(tl):
    """VABSDIFF4 warp-instructions one ME launch of the current configuration cannot avoid (two lists): HME level 0 (16 x 8-row block, every position of the
    search area: one (64 x 32) * mult quadrant at M10 4K, 2 x 2 regions of (32 x 16) * mult at M7), [M7: HME level 1, 4 regions of 8 x 4 on a 32 x 16-row block],
    the full-pel search (8 x 7 or 16 x 9 positions on 64 x 32 source rows), six + one LCU-sized centre tests, bi-prediction.  Sub-pel refinement is not counted."""
    if ENC_MODE == 7:
        hme = 4 * ((32 * HME_MULT[tl] // 100) * (16 * HME_MULT[tl] // 100)) * 32 + 4 * 32 * 128
        fp = 144 * 512
    else:
        saw, sah = (64 * HME_MULT[tl] // 100) & ~15, 32 * HME_MULT[tl] // 100
        hme, fp = saw * sah * 32, 56 * 512
    per_list = hme + fp + 7 * 512
    bipred = 3 * 4096 // 2 // 4                         # 64x64 + 4 32x32 + 16 16x16 PUs, every other row, 4 samples per lane-op
    last_full = (H % 64) == 0
    full_rows = (ROWS if last_full else ROWS - 1) * COLS          # HME runs on full-height LCUs only
    part = 0 if last_full else (2 * (fp + 7 * 384) + bipred) * COLS
    return ((2 * per_list + bipred) * full_rows + part) / 32.0


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
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
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
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# the workload: synthetic frames + the fixed decision tables shared by both arms
# ------------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, n_frames, seed):
        import common
        import svtb200 as S
        self.S = S
        if W > 4096:      # 8K: a 2 x 2 tiling of a 4K clip (same texture statistics and motion, a quarter of the generation time)
            v = common.SynthVideo(W // 2, H // 2, seed=seed)
            self.luma = [np.tile(v.luma(i), (2, 2)) for i in range(n_frames)]
            self.chroma = [tuple(np.tile(c, (2, 2)) for c in v.chroma(i)) for i in range(n_frames)]
        else:
            v = common.SynthVideo(W, H, seed=seed)
            self.luma = [v.luma(i) for i in range(n_frames)]
            self.chroma = [v.chroma(i) for i in range(n_frames)]
        self.n = n_frames
        rng = np.random.default_rng(seed ^ 0xABCD)
        if BITS == 10:
            # 10-bit = the 8-bit clip x 4 + 2 random LSBs, held as the encoder holds its input: 8-bit MSB planes + 2-bit planes (byte = lsb << 6)
            self.lsb = [tuple(((rng.integers(0, 4, p.shape, dtype=np.uint8)) << 6).astype(np.uint8) for p in (self.luma[i],) + tuple(self.chroma[i]))
                        for i in range(n_frames)]
        # geometry of the last LCU row (4K: 48 samples, 8K: 32, 1080p: 56): the largest of 32 / 16 / 8 that tiles it
        self.last_h = H - 64 * (ROWS - 1)
        self.n_last = 32 if self.last_h % 32 == 0 else 16 if self.last_h % 16 == 0 else 8
        # TU jobs: 32x32 luma + 16x16 chroma; n_last (/ 2) in the bottom LCUs
        jobs = []
        for plane, sh in ((0, 0), (1, 1), (2, 1)):
            for ly in range(ROWS):
                n = (32 >> sh) if ly < ROWS - 1 else (self.n_last >> sh)
                lh = (64 if ly < ROWS - 1 else self.last_h) >> sh
                for lx in range(COLS):
                    for yy in range(0, lh, n):
                        for xx in range(0, 64 >> sh, n):
                            jobs.append(((lx * 64 >> sh) + xx, (ly * 64 >> sh) + yy, plane, n))
        tq = np.zeros(len(jobs), dtype=S.TQ_JOB_DTYPE)
        arr = np.array(jobs)
        tq["x"], tq["y"], tq["plane"], tq["size"] = arr[:, 0], arr[:, 1], arr[:, 2], arr[:, 3]
        tq["qp"] = np.where(tq["plane"] == 0, 32, 31)
        tq["slice_type"] = 0
        self.tq_jobs = tq
        # boundary strengths: 1 on every PU/TU boundary (32-sample grid; n_last grid in the bottom row), none on the picture border
        vbs = np.zeros((ROWS, COLS, 16, 16), np.uint8)
        hbs = np.zeros((ROWS, COLS, 16, 16), np.uint8)
        vbs[:, :, :, 0] = 1; vbs[:, :, :, 8] = 1; hbs[:, :, 0, :] = 1; hbs[:, :, 8, :] = 1
        for k in range(0, 16, self.n_last // 4):
            vbs[-1, :, :, k] = 1
            if 4 * k < self.last_h:
                hbs[-1, :, k, :] = 1
        vbs[:, 0, :, 0] = 0; hbs[0, :, 0, :] = 0
        self.vbs, self.hbs = vbs.reshape(N_LCU, 256), hbs.reshape(N_LCU, 256)
        self.qp_array = np.full(((H + 7) // 8, W // 8), 32, np.uint8)
        self.dlf_prm = S.DlfParams(W, H, W // 8, 0, 0, 0, 0)
        prm = np.zeros(N_LCU, dtype=S.SAO_PARAMS_DTYPE)
        prm["type"] = rng.integers(0, 6, (N_LCU, 2))
        prm["band_position"] = rng.integers(0, 29, (N_LCU, 3))
        prm["offset"] = rng.integers(-3, 4, (N_LCU, 3, 4))
        self.sao_prm = prm

    def tl(self, step):
        return TL_PATTERN[step % 8]

    def me_cfg(self, step):
        return self.me_cfg_tl(self.tl(step))

    def me_cfg_tl(self, tl):
        return self.S.me_config(W, H, enc_mode=ENC_MODE, temporal_layer=tl, num_lists=2, ref_poc=(0, 16), frame_rate=FPS)

    def ois_cfg(self, tl=1):
        """non-I pictures (EbPictureDecisionProcess.c:417-455): M10 DC only; M7 the full open-loop search"""
        return self.S.ois_config(W, H, enc_mode=ENC_MODE, temporal_layer=tl, slice_type="B")

    def ois_cap(self):
        return 1 if ENC_MODE >= 10 else 9

    def refs(self, step):
        d = {0: 8, 1: 4, 2: 2, 3: 1}[self.tl(step)]
        return (step - d) % self.n, (step + d) % self.n

    def mc_jobs(self, me_results):
        """Mode-decision stand-in: each 32x32 PU (n_last x n_last in the bottom LCU row) takes its best ME candidate."""
        S = self.S
        res = me_results.reshape(ROWS, COLS, 85)
        out = []
        for ly in range(ROWS):
            if ly < ROWS - 1:
                pus = [(1 + k, (k & 1) * 32, (k >> 1) * 32, 32) for k in range(4)]
            elif self.n_last == 32:
                pus = [(1 + k, (k & 1) * 32, (k >> 1) * 32, 32) for k in range(4) if (k >> 1) * 32 < self.last_h]
            elif self.n_last == 16:
                pus = [(5 + by * 4 + bx, bx * 16, by * 16, 16) for by in range(self.last_h // 16) for bx in range(4)]
            else:
                pus = [(21 + by * 8 + bx, bx * 8, by * 8, 8) for by in range(self.last_h // 8) for bx in range(8)]
            for (pu, px, py, sz) in pus:
                r = res[ly, :, pu]
                j = np.zeros(COLS, dtype=S.MC_JOB_DTYPE)
                j["x"] = np.arange(COLS) * 64 + px
                j["y"] = ly * 64 + py
                j["w"] = j["h"] = sz
                j["dir"] = r["direction"][:, 0]
                j["ref_idx"][:, 0], j["ref_idx"][:, 1] = 0, 1
                j["mv"][:, 0, 0], j["mv"][:, 0, 1] = r["x_mv_l0"], r["y_mv_l0"]
                j["mv"][:, 1, 0], j["mv"][:, 1, 1] = r["x_mv_l1"], r["y_mv_l1"]
                out.append(j)
        return np.concatenate(out)

    @staticmethod
    def mc_alg_bytes(jobs):
        px = jobs["w"].astype(np.int64) * jobs["h"] * 3 // 2
        return int((px * (np.where(jobs["dir"] == 2, 2, 1) + 1)).sum()) * (2 if BITS == 10 else 1)

    def luma16(self, i):
        return (self.luma[i].astype(np.uint16) << 2) | (self.lsb[i][0] >> 6)

    def chroma16(self, i):
        return tuple((self.chroma[i][k].astype(np.uint16) << 2) | (self.lsb[i][1 + k] >> 6) for k in range(2))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import svtb200 as S

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    cfg = CONFIGS[CFG_ID]
    numa = bind_to_gpu_numa_node(local)      # before anything is pinned: first touch of the pinned buffers lands on the GPU's NUMA node
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S.check(S.lib().svtb200_init(C.c_int(local)))
    tstream = torch.cuda.current_stream()
    stream = tstream.cuda_stream
    wl = Workload(RING, seed=0x5F7C0001 + rank)
    dev = torch.device("cuda", local)
    hbd = BITS == 10
    bps = 2 if hbd else 1
    sdt = torch.int16 if hbd else torch.uint8           # sample storage (16-bit samples are unsigned; torch only carries the bits)
    alg_ctu = alg_per_ctu()
    # 8-bit or 16-bit flavour of every EncDec-side entry point
    F = dict(mc=S.mc_batch16_dev if hbd else S.mc_batch_dev, tq=S.tq_batch16_dev if hbd else S.tq_batch_dev, dlf=S.dlf_picture16_dev if hbd else S.dlf_picture_dev,
             sao_stats=S.sao_stats_picture16_dev if hbd else S.sao_stats_picture_dev, sao_h3=S.sao_stats_picture16_h3_dev if hbd else S.sao_stats_picture_h3_dev,
             sao_apply=S.sao_apply_picture16_dev if hbd else S.sao_apply_picture_dev)

    def dev_yuv(dtype=None):
        dtype = dtype or sdt
        y = torch.zeros((H + 2 * PAD, W + 2 * PAD), dtype=dtype, device=dev)
        cb = torch.zeros((H // 2 + PAD, W // 2 + PAD), dtype=dtype, device=dev)
        cr = torch.zeros((H // 2 + PAD, W // 2 + PAD), dtype=dtype, device=dev)
        st = S.make_yuv_struct((y.data_ptr(), cb.data_ptr(), cr.data_ptr()), (y.shape[1], cb.shape[1], cr.shape[1]), PAD, W, H)
        return (y, cb, cr), st

    def fill_yuv(t, y, cb, cr, replicate):
        """H2D of the three planes into the padded device picture (+ edge replication when it will serve as a reference)."""
        t[0][PAD:PAD + H, PAD:PAD + W].copy_(y, non_blocking=True)
        t[1][PAD // 2:PAD // 2 + H // 2, PAD // 2:PAD // 2 + W // 2].copy_(cb, non_blocking=True)
        t[2][PAD // 2:PAD // 2 + H // 2, PAD // 2:PAD // 2 + W // 2].copy_(cr, non_blocking=True)
        if replicate:
            for pl, p in zip(t, (PAD, PAD // 2, PAD // 2)):
                hh, ww = pl.shape[0] - 2 * p, pl.shape[1] - 2 * p
                pl[p:p + hh, :p] = pl[p:p + hh, p:p + 1]; pl[p:p + hh, p + ww:] = pl[p:p + hh, p + ww - 1:p + ww]
                pl[:p, :] = pl[p:p + 1, :]; pl[p + hh:, :] = pl[p + hh - 1:p + hh, :]

    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_y = [pin(a) for a in wl.luma]                     # 8-bit luma: what PictureAnalysis / ME / OIS read at any bit depth
    if hbd:
        # e2e input = unpacked 10-bit planes (yuv420p10le); resident input = the encoder's own 8-bit + 2-bit input planes, packed on the device per picture
        h_s = [tuple(pin(a.view(np.int16)) for a in (wl.luma16(i),) + wl.chroma16(i)) for i in range(RING)]
        d_in8 = [tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (wl.luma[i],) + tuple(wl.chroma[i])) for i in range(RING)]
        d_inn = [tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in wl.lsb[i]) for i in range(RING)]
    else:
        h_s = [(h_y[i], pin(wl.chroma[i][0]), pin(wl.chroma[i][1])) for i in range(RING)]
    ring_pa = [S.DevPaPicture(W, H) for _ in range(RING)]
    ring_yuv = [dev_yuv() for _ in range(RING)]          # reconstructed references (16-bit samples at 10 bits) == the sources of the synthetic clip
    for i in range(RING):
        ring_pa[i].upload(h_y[i].data_ptr(), stream)
        fill_yuv(ring_yuv[i][0], h_s[i][0], h_s[i][1], h_s[i][2], True)

    cap_ois = wl.ois_cap()
    n_tq = len(wl.tq_jobs)
    stream_cap = int(sum(int(j["size"]) ** 2 for j in wl.tq_jobs))       # elements of the packed coefficient stream when every TU is coded
    # what the host consumes of a picture: ME records, TU offsets, nz counts, SAO decisions (the statistics stay on the device: the RD decision runs there),
    # OIS totals and candidates
    small_sizes = [N_LCU * 85 * 24, (n_tq + 1) * 4, n_tq * 4, N_LCU * 40, N_LCU * 85, N_LCU * 85 * cap_ois * 4]
    small_offs = np.concatenate([[0], np.cumsum([(b + 255) // 256 * 256 for b in small_sizes])]).astype(int)
    small_total = int(small_offs[-1])
    frame_capacity = small_total + 2 * stream_cap        # [small results | packed coefficient stream]: one contiguous span per picture

    class Slot:
        """device + pinned host buffers of one frame in flight (the e2e loop keeps several in flight: copies of one overlap the kernels of the other)"""
        def __init__(self, own_outputs=True):
            self.pred_t, self.pred = dev_yuv()
            self.sao_t, self.sao_out = dev_yuv()
            self.unf_t, self.unf = dev_yuv()                 # reconstruction before deblocking (H3 statistics of layer-0 pictures)
            self.stats_t = torch.zeros(N_LCU * 3 * 312, dtype=torch.uint8, device=dev)       # SAO statistics: consumed by the decision on the device
            self.sao_prm_t = torch.zeros(N_LCU * 20, dtype=torch.uint8, device=dev)          # the decision in the form sao_apply reads
            self.in_t, self.in_yuv = dev_yuv()
            if hbd:
                self.in8 = torch.zeros((H, W), dtype=torch.uint8, device=dev)      # e2e: 8-bit luma of the uploaded 10-bit picture (UnPack8BitData)
                self.src_t, self.src = dev_yuv()                                   # resident: the packed 16-bit source picture
            self.pa = S.DevPaPicture(W, H)
            self.coeff_t = [torch.zeros((H, W), dtype=torch.int16, device=dev), torch.zeros((H // 2, W // 2), dtype=torch.int16, device=dev),
                            torch.zeros((H // 2, W // 2), dtype=torch.int16, device=dev)]
            self.coeff = S.CoeffPicture()
            for k in range(3):
                self.coeff.buf[k] = self.coeff_t[k].data_ptr(); self.coeff.stride[k] = self.coeff_t[k].shape[1]
            self.cap = stream_cap
            self.small_bytes = int(sum(small_sizes))
            self.pending = False
            self.ev_small = torch.cuda.Event()
            self.ev_in, self.ev_cmp, self.ev_out = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
            self.ev_dec, self.dec_pending = torch.cuda.Event(), False      # the SAO decision of the slot's last picture, when it ran on the side stream
            if own_outputs:
                # everything the host consumes sits in ONE device span per picture (small results, then the packed stream): ME results, TU offsets, nz counts,
                # SAO stats, OIS; the coefficients leave as a dense stream of the coded TUs (svtb200_tq_pack_coeffs_dev)
                self.d_out = torch.zeros(frame_capacity, dtype=torch.uint8, device=dev)
                self.h_small = torch.zeros(small_total, dtype=torch.uint8).pin_memory()
                self.h_stream = torch.empty(stream_cap, dtype=torch.int16).pin_memory()
                self.h_off_total = self.h_small[small_offs[1]:small_offs[1] + small_sizes[1]].view(torch.int32).numpy()      # [-1] = packed stream length
                self.bind(self.d_out)

        def bind(self, span):
            """point the per-picture outputs at `span` (frame_capacity bytes of device memory: this slot's own, or a picture's place in the GOP-exchange buffer)"""
            base = span.data_ptr()
            o = small_offs
            self.p = dict(res=base + int(o[0]), off=base + int(o[1]), nz=base + int(o[2]), dec=base + int(o[3]), stats=self.stats_t.data_ptr(), sao_prm=self.sao_prm_t.data_ptr(),
                          ois_t=base + int(o[4]), ois_c=base + int(o[5]),
                          small=base, stream=base + small_total, yint=self.in_t[0][PAD:, PAD:].data_ptr())
            if hasattr(self, "h_small"):
                self.p.update(h_small=self.h_small.data_ptr(), h_stream=self.h_stream.data_ptr())
            self.off_total = span[int(o[1]) + 4 * n_tq:int(o[1]) + 4 * n_tq + 4].view(torch.int32)

    NS = max(3, args.inflight)            # e2e: NS frames in flight (H2D / kernels / D2H of consecutive frames overlap, the host runs NS - 1 frames ahead)
    if W > 4096:
        NS = 3                            # 8K: 0.8 GB of device pictures per slot
    slots = [Slot() for _ in range(NS)]
    d_tq = torch.from_numpy(wl.tq_jobs.view(np.uint8)).to(dev)
    d_vbs, d_hbs, d_qp = (torch.from_numpy(a.copy()).to(dev) for a in (wl.vbs, wl.hbs, wl.qp_array))
    d_sao_prm = torch.from_numpy(wl.sao_prm.view(np.uint8)).to(dev)
    ois_cfgs = {tl: wl.ois_cfg(tl) for tl in set(TL_PATTERN)}

    # mode-decision stand-in: one ME pass per (ring frame, layer) pair fixes the MC job lists used by both timed loops
    d_mc, mc_bytes = [], []
    d_res0 = torch.zeros(small_sizes[0], dtype=torch.uint8, device=dev)
    for i in range(RING):
        r0, r1 = wl.refs(i)
        S.me_picture_dev(wl.me_cfg(i), ring_pa[i], ring_pa[r0], ring_pa[r1], d_res0.data_ptr(), stream)
        torch.cuda.synchronize()
        jobs = wl.mc_jobs(d_res0.cpu().numpy().view(S.ME_PU_RESULT_DTYPE).reshape(N_LCU, 85))
        d_mc.append((torch.from_numpy(jobs.view(np.uint8)).to(dev), len(jobs)))
        mc_bytes.append(wl.mc_alg_bytes(jobs))
    torch.cuda.synchronize()

    names = (["pack"] if hbd else []) + ["me", "ois", "mc", "tq", "h3_copy", "dlf", "sao_stats", "sao_dec", "sao_apply"] + (["pack_coeffs"] if cfg["gop_exchange"] else [])
    # per-frame constants hoisted out of the loops (the loops should cost the host as little as a C caller would spend)
    me_cfgs = {tl: wl.me_cfg_tl(tl) for tl in set(TL_PATTERN)}
    ref_pairs = {}
    for i in range(RING):
        r0, r1 = wl.refs(i)
        ref_pairs[(r0, r1)] = S.yuv_struct_array([ring_yuv[r0][1], ring_yuv[r1][1]])
    d_mc_ptr = [(t.data_ptr(), n) for t, n in d_mc]
    p_tq, p_vbs, p_hbs, p_qp, p_sao_prm = d_tq.data_ptr(), d_vbs.data_ptr(), d_hbs.data_ptr(), d_qp.data_ptr(), d_sao_prm.data_ptr()
    rates_b = sao_rates()
    hp = [tuple(t.data_ptr() for t in h_s[k]) for k in range(RING)]

    def pack_source(k, sl, st):
        """10 bits: the encoder's 8-bit + 2-bit input planes -> the 16-bit source picture of the EncDec stages (Pack2D_SRC, EbPictureOperators.c:544), per plane"""
        for pl, (w_, h_, pad) in enumerate(((W, H, PAD), (W // 2, H // 2, PAD // 2), (W // 2, H // 2, PAD // 2))):
            S.pack_plane_dev(S.PACK_MSB, d_in8[k][pl].data_ptr(), w_, d_inn[k][pl].data_ptr(), w_, sl.src_t[pl][pad:, pad:].data_ptr(), sl.src_t[pl].shape[1], 0, 0, w_, h_, st)

    def side_decision(sl, ts, timed=None):
        """No picture stage of a frame outside layer 0 waits for its SAO decision (SAO is not applied there): the decision's dependency chain (an anti-diagonal
        wavefront in one CTA) runs on a side stream next to the following frame's kernels; whoever consumes the decisions (the copy-out, the GOP exchange)
        waits for ev_dec."""
        P, side = sl.p, dstreams[ts]
        e = torch.cuda.Event(); e.record(ts); side.wait_event(e)
        fn = lambda: S.sao_decision_picture_dev(P["stats"], COLS, ROWS, rates_b, SAO_LAMBDA[0], SAO_LAMBDA[1], 1 if hbd else 0, P["dec"], P["sao_prm"], side.cuda_stream)
        if timed is not None:
            timed("sao_dec", fn, on=side)
        else:
            fn()
        sl.ev_dec.record(side); sl.dec_pending = True

    def pipeline(i, src_pa, src_yuv, sl, ts, ev=None, resident=True, inline_dec=False):
        """All device stages of frame i into slot sl on torch stream ts.  ev: dict name -> list of (start, end) events."""
        st = ts.cuda_stream
        def timed(name, fn, on=None):
            if ev is None:
                fn(); return
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(on or ts); fn(); b.record(on or ts)
            ev[name].append((a, b))
        r0, r1 = wl.refs(i)
        k = i % RING
        tl = wl.tl(i)
        tl0 = tl == 0
        pred = sl.pred
        P = sl.p
        if hbd and resident:
            timed("pack", lambda: pack_source(k, sl, st))
            src_yuv = sl.src
        timed("me", lambda: S.me_picture_dev(me_cfgs[tl], src_pa, ring_pa[r0], ring_pa[r1], P["res"], st))
        timed("ois", lambda: S.ois_picture_dev(ois_cfgs[tl], src_pa.struct.full, P["res"], P["ois_t"], P["ois_c"], cap_ois, st))
        timed("mc", lambda: F["mc"](ref_pairs[(r0, r1)], d_mc_ptr[k][0], d_mc_ptr[k][1], pred, st))
        timed("tq", lambda: F["tq"](src_yuv, pred, pred, sl.coeff, p_tq, n_tq, P["nz"], st))
        if sl.dec_pending:                      # the slot's statistics buffer may still be read by its previous picture's decision (side stream): the writer waits
            ts.wait_event(sl.ev_dec); sl.dec_pending = False
        if tl0:
            # H3 (EbCodingLoop.c:4604-4635, 4711-4748): the encoder gathers LCU n's statistics right after LCU n's own deblocking step, i.e. from the
            # deblocked picture except the right / bottom 4-sample strips of every LCU, which are still unfiltered: keep the reconstruction before DLF
            timed("h3_copy", lambda: S.yuv_picture_copy_dev(pred, sl.unf, st, bps))
            timed("dlf", lambda: F["dlf"](pred, p_vbs, p_hbs, p_qp, wl.dlf_prm, st))
            timed("sao_stats", lambda: F["sao_h3"](src_yuv, pred, sl.unf, P["stats"], st))
            timed("sao_dec", lambda: S.sao_decision_picture_dev(P["stats"], COLS, ROWS, rates_b, SAO_LAMBDA[0], SAO_LAMBDA[1], 1 if hbd else 0, P["dec"], P["sao_prm"], st))
            timed("sao_apply", lambda: F["sao_apply"](pred, sl.sao_out, P["sao_prm"], (1, 1), st))
        else:
            timed("sao_stats", lambda: F["sao_stats"](src_yuv, pred, P["stats"], st))
            if inline_dec:      # (a captured frame graph is one stream: the caller runs this frame's decision after the replay, see side_decision)
                return
            side_decision(sl, ts, timed)
        if cfg["gop_exchange"] and resident:
            timed("pack_coeffs", lambda: S.tq_pack_coeffs_dev(sl.coeff, p_tq, P["nz"], n_tq, P["stream"], sl.cap, P["off"], st))

    # resident loop: frames are independent, so consecutive frames alternate between `--streams` compute streams (one output slot each);
    # kernels of different frames then share the SMs, which hides launch tails and lets the latency-bound kernels overlap the issue-bound ones.
    # config 5: the compute streams ARE the concurrent video streams of the density run (frame i belongs to video stream i % streams).
    cstreams = [tstream] + [torch.cuda.Stream(device=dev) for _ in range(args.streams - 1)]
    dstreams = {cs: torch.cuda.Stream(device=dev) for cs in cstreams}      # one side stream per compute stream: SAO decisions off the critical path
    while len(slots) < len(cstreams):
        slots.append(Slot())

    # config 4: closed GOPs round-robin over the ranks; the pictures of a GOP write their outputs straight into the GOP-exchange buffer
    FPG = 16
    ex = None
    if cfg["gop_exchange"]:
        from gop_exchange import GopExchange
        ex = GopExchange(rank, world, dev, FPG, frame_capacity, peer=(args.exchange == "peer"))
        comm_stream = torch.cuda.Stream(device=dev)
        ex_stats = {"rounds": 0, "bytes": 0, "ms": []}

    ev_round = [None, None]              # per parity: the transfers that last read this half of the exchange buffer

    def step_resident(i, ev=None, frame_in_gop=None, parity=0):
        k = i % len(cstreams)
        sl = slots[k]
        if ex is not None:
            if frame_in_gop == 0 and ev_round[parity] is not None:
                for cs in cstreams:
                    cs.wait_event(ev_round[parity])                                          # GOP k - 2 has left this half of the buffer
            sl.bind(ex.frame_view(parity, frame_in_gop))
        pipeline(i, ring_pa[i % RING], ring_yuv[i % RING][1], sl, cstreams[k], ev)

    def gop_boundary(round_index, gop_slots_totals):
        """all pictures of this rank's GOP are done: their byte counts reach the host (one small D2H), the transfers to rank 0 are posted on the
        communication stream and overlap the next GOP's kernels"""
        for cs in list(cstreams) + list(dstreams.values()):
            e = torch.cuda.Event(); e.record(cs); comm_stream.wait_event(e)
        with torch.cuda.stream(comm_stream):
            totals = torch.stack(gop_slots_totals).flatten().cpu().tolist()                 # elements of each picture's packed stream
            fb = [small_total + 2 * min(int(v), stream_cap) for v in totals]
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(comm_stream)
            moved = ex.exchange(round_index, fb)
            ex.wait_all()                                                                    # stream-ordered: the communication stream waits, the host does not
            b.record(comm_stream)
            ev_round[round_index & 1] = b
        ex_stats["rounds"] += 1; ex_stats["bytes"] += moved; ex_stats["ms"].append((a, b))

    # copy-in, small-results copy-out and coefficient-stream copy-out each get a stream: a frame's stream copy is issued two frames later (its
    # length must reach the host first) and must not queue behind the next frame's kernels
    s_in, s_out, s_small = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    d2h_bytes = []
    skip = set(x for x in args.e2e_skip.split(",") if x)

    def finish_slot(sl):
        """second half of a frame's D2H: the packed coefficient stream, whose length is known once the offsets have arrived"""
        if not sl.pending:
            return
        sl.ev_small.synchronize()
        total = min(int(sl.h_off_total[-1]), sl.cap)
        if "d2h" not in skip:
            S.copy_async(sl.p["h_stream"], sl.p["stream"], 2 * total, True, s_out.cuda_stream)
        sl.ev_out.record(s_out)
        d2h_bytes.append(2 * total + sl.small_bytes)
        sl.pending = False

    def frame_kernels(i, sl, cs, inline_dec):
        """the device stages of e2e frame i on compute stream cs"""
        if "pa" not in skip:
            if hbd:      # UnPack8BitData (EbPackUnPack_C.c:153): the 8-bit luma PictureAnalysis / ME / OIS work on
                S.pack_plane_dev(S.PACK_UNPACK8, sl.p["yint"], sl.in_t[0].shape[1], 0, 0, sl.in8.data_ptr(), W, 0, 0, W, H, cs.cuda_stream)
                sl.pa.build_from_device(sl.in8.data_ptr(), W, cs.cuda_stream)
            else:
                sl.pa.build_from_device(sl.p["yint"], sl.in_t[0].shape[1], cs.cuda_stream)     # pad + decimate on the device
        pipeline(i, sl.pa, sl.in_yuv, sl, cs, resident=False, inline_dec=inline_dec)
        if "pack" not in skip:
            S.tq_pack_coeffs_dev(sl.coeff, p_tq, sl.p["nz"], n_tq, sl.p["stream"], sl.cap, sl.p["off"], cs.cuda_stream)

    # The e2e loop is bound by the HOST when every stage is a separate call (about 20 library calls and as many stream / event operations per frame from
    # Python): the frame's kernels are captured once per (slot, ring picture, stream) combination into a CUDA graph and replayed -- one launch per frame.
    # Pointers and per-layer parameters are baked in, so the combinations must repeat: 4 slots and 4 streams against the 16-picture ring = 16 graphs.
    GRAPH_PERIOD = RING
    frame_graphs = {}
    e2e_streams = cstreams
    if not args.no_graphs and RING % 4 == 0:
        NS = 4
        while len(slots) < NS:
            slots.append(Slot())
        e2e_streams = [torch.cuda.Stream(device=dev) for _ in range(4)]      # (capture needs non-default streams)
        for cs in e2e_streams:
            dstreams[cs] = torch.cuda.Stream(device=dev)

    def capture_frame_graphs():
        for i in range(GRAPH_PERIOD):
            sl, cs = slots[i % NS], e2e_streams[i % len(e2e_streams)]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=cs):
                frame_kernels(i, sl, cs, inline_dec=True)
            frame_graphs[i] = g

    def step_e2e(i):
        """Frame i from pinned host buffers: H2D on the copy-in stream, kernels on the compute stream, D2H of everything the host-side
        encoder consumes on the copy-out stream; NS frames in flight, every copy inside the timed region."""
        k = i % RING
        sl = slots[i % NS]
        finish_slot(sl)                                                                    # frame i - NS (same slot) is completely out
        s_in.wait_event(sl.ev_cmp)                                                         # the kernels that read this slot's inputs are done
        if "h2d" not in skip:
            S.yuv_picture_upload(sl.in_yuv, hp[k][0], W, hp[k][1], W // 2, hp[k][2], W // 2, bps, s_in.cuda_stream)    # H2D: Y + Cb + Cr (2 bytes / sample at 10 bits)
        sl.ev_in.record(s_in)
        cs = e2e_streams[i % len(e2e_streams)]                                             # consecutive frames alternate compute streams, as in the resident loop
        cs.wait_event(sl.ev_in)
        cs.wait_event(sl.ev_out)                                                           # this slot's previous results have left the device
        g = frame_graphs.get(i % GRAPH_PERIOD) if frame_graphs else None
        if g is not None:
            if sl.dec_pending:                                                             # (the wait the call-by-call pipeline issues before its statistics kernel)
                cs.wait_event(sl.ev_dec); sl.dec_pending = False
            with torch.cuda.stream(cs):
                g.replay()                                                                 # the frame's kernels as one captured graph: one launch from the host
            if wl.tl(i) != 0:
                side_decision(sl, cs)
        else:
            frame_kernels(i, sl, cs, inline_dec=False)
        sl.ev_cmp.record(cs)
        s_small.wait_event(sl.ev_cmp)
        if sl.dec_pending:
            s_small.wait_event(sl.ev_dec)
        if "d2h" not in skip:
            S.copy_async(sl.p["h_small"], sl.p["small"], small_total, True, s_small.cuda_stream)    # D2H: ME results, TU offsets, nz, SAO decisions, OIS
        sl.ev_small.record(s_small)
        sl.pending = True
        finish_slot(slots[(i + 1) % NS])                                                   # frame i - NS + 1: its offsets arrived long ago, its stream goes out now

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if ex is not None:
        args.steps = max(FPG, args.steps // FPG * FPG)                                     # whole GOPs
    for i in range(args.warmup):
        step_resident(i, frame_in_gop=i % FPG, parity=0)
    sampler = ClockSampler(local); sampler.start()
    # the timed region (exactly --steps frames between two barriers) is run --repeats times; the MEDIAN run is the one reported
    runs = []
    round_counter = 0
    for rep in range(max(1, args.repeats)):
        launches0 = S.lib().svtb200_launch_count()
        ev = {n: [] for n in names}
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0.record(tstream)
        for cs in cstreams[1:]:
            cs.wait_event(t0)
        totals = []
        for i in range(args.steps):
            step_resident(args.warmup + i, ev, frame_in_gop=i % FPG, parity=round_counter & 1)
            if ex is not None:
                totals.append(slots[(args.warmup + i) % len(cstreams)].off_total)
                if (i + 1) % FPG == 0:
                    gop_boundary(round_counter, totals); round_counter += 1; totals = []
        for cs in cstreams[1:]:
            e = torch.cuda.Event(); e.record(cs); tstream.wait_event(e)
        if ex is not None:
            ex.wait_all()
            e = torch.cuda.Event(); e.record(comm_stream); tstream.wait_event(e)
        t1.record(tstream)
        barrier()
        runs.append((t0.elapsed_time(t1), {n: [a.elapsed_time(b) for a, b in ev[n]] for n in names}, S.lib().svtb200_launch_count() - launches0))
    sampler.stop_flag = True
    ex_report = None
    if ex is not None:
        # what moved, seen from the producers (in peer mode rank 0 runs no transfer of its own): per rank the bytes and the time its copies spent on the
        # communication stream, and a check that rank 0 holds what the producers sent (byte sum of the first picture of every rank's last GOP)
        mine = {"bytes": int(ex_stats["bytes"]), "ms": [a.elapsed_time(b) for a, b in ex_stats["ms"]]}
        last_parity = (round_counter - 1) & 1
        mine["sum0"] = int(ex.send[last_parity][:small_total].to(torch.int64).sum().item())
        allr = [None] * world
        if world > 1:
            dist.all_gather_object(allr, mine)
        else:
            allr = [mine]
        if rank == 0:
            ok = True
            for src in range(1, world):
                got = ex.payload_on_rank0((round_counter - 1) * world + src, 0)[:small_total].to(torch.int64).sum().item()
                ok = ok and int(got) == allr[src]["sum0"]
            ex_report = {"per_rank": allr, "payload_verified": bool(ok)}
        for sl in slots:
            sl.bind(sl.d_out)                                                               # the e2e loop below uses the slots' own spans again
    runs_ms = [r[0] for r in runs]
    ms, stage_ms, launches = sorted(runs, key=lambda r: r[0])[len(runs) // 2]
    # the same stages with nothing else on the GPU (one stream): what each kernel costs alone -- reported next to the timed-region numbers
    ev1 = {n: [] for n in names}
    for i in range(min(args.steps, 16)):
        pipeline(args.warmup + i, ring_pa[(args.warmup + i) % RING], ring_yuv[(args.warmup + i) % RING][1], slots[0], tstream, ev1)
    barrier()
    serial_ms = {n: [a.elapsed_time(b) for a, b in ev1[n]] for n in names}

    # what the host link gives this process (pinned buffers, one direction at a time and both at once): the e2e figure is read against it
    def pcie_probe(nbytes=64 << 20, reps=4):
        hb, hb2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        db, db2 = torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.empty(nbytes, dtype=torch.uint8, device=dev)
        res = {}
        for name, fn in (("h2d_gbs", lambda: db.copy_(hb, non_blocking=True)), ("d2h_gbs", lambda: hb2.copy_(db2, non_blocking=True))):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            res[name] = nbytes * reps / (a.elapsed_time(b) * 1e-3) / 1e9
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s_in):
                db.copy_(hb, non_blocking=True)
            with torch.cuda.stream(s_out):
                hb2.copy_(db2, non_blocking=True)
        torch.cuda.synchronize()
        res["duplex_each_gbs"] = nbytes * reps / (time.perf_counter() - t0) / 1e9
        return res
    if world > 1:
        dist.barrier()                     # every rank probes at the same time: what the link gives when all GPUs of the node copy at once
    pcie = pcie_probe()

    for i in range(args.warmup):
        step_e2e(i)
    for sl in slots[:NS]:
        finish_slot(sl)
    barrier()
    graph_note = "off (--no-graphs)"
    if not args.no_graphs and RING % 4 == 0:
        try:
            for sl in slots:
                sl.dec_pending = False      # (everything has completed at the barrier: no event of the call-by-call frames may enter a capture)
            capture_frame_graphs()
            barrier()
            for i in range(GRAPH_PERIOD):      # first replay of every graph outside the timed region
                step_e2e(args.warmup + i)
            for sl in slots[:NS]:
                finish_slot(sl)
            barrier()
            graph_note = f"{len(frame_graphs)} frame graphs (pa build + pipeline + coefficient pack per frame), {NS} slots, {len(e2e_streams)} streams"
        except Exception as exc:      # capture refused on this driver / torch: fall back to call-by-call launches and say so
            frame_graphs.clear()
            torch.cuda.synchronize()
            graph_note = f"capture failed, call-by-call launches: {type(exc).__name__}: {str(exc)[:160]}"
    w0 = time.perf_counter()
    d2h_bytes.clear()
    for i in range(args.steps):
        step_e2e(args.warmup + i)
    for q in range(1, NS + 1):
        finish_slot(slots[(args.warmup + args.steps - 1 + q) % NS])
    barrier()
    e2e_s = time.perf_counter() - w0

    tms = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    link = torch.tensor([pcie["h2d_gbs"], pcie["d2h_gbs"], pcie["duplex_each_gbs"]], dtype=torch.float64, device="cuda")
    link_all = [link.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_gather(link_all, link)
    ms, e2e_ms = float(tms[0]), float(tms[1])
    if rank == 0:
        peak, peak_src = peaks()
        ctus = N_LCU * args.steps * world
        tot = {n: float(np.sum(v)) for n, v in stage_ms.items()}
        # the dominant kernel by SM time (the device-to-device copy and the SAO decision -- one CTA walking a dependency chain on a side stream -- are not candidates)
        dom = max((n for n in tot if n not in ("h3_copy", "sao_dec")), key=tot.get)
        alg = {"me": alg_ctu["me"] * N_LCU, "ois": alg_ctu["ois"] * N_LCU, "mc": float(np.mean(mc_bytes)), "tq": alg_ctu["tq"] * N_LCU, "dlf": alg_ctu["dlf"] * N_LCU,
               "h3_copy": 2 * 6144 * bps * N_LCU, "sao_stats": alg_ctu["sao_stats"] * N_LCU, "sao_dec": ALG["sao_dec"] * N_LCU, "sao_apply": alg_ctu["sao_apply"] * N_LCU, "pack": alg_ctu["pack"] * N_LCU,
               "pack_coeffs": 2 * 12288 * N_LCU}
        per_stage = {n: {"launch_ms_avg": float(np.mean(v)) if v else None, "launches": len(v), "share_of_step": tot[n] / sum(tot.values()),
                         "achieved_gbs": (alg[n] / (np.mean(v) * 1e-3) / 1e9) if v else None,
                         "frac_of_hbm_peak": (alg[n] / (np.mean(v) * 1e-3) / 1e9 / peak) if v else None} for n, v in stage_ms.items()}
        k_avg_s = float(np.mean(stage_ms[dom])) * 1e-3
        traffic, traffic_src = ncu_traffic() if CFG_ID == 2 else ({}, None)
        sm_clk = (sampler.summary().get("sm_mhz") or 1965.0) * 1e6
        sad_peak = 148 * 4 * VABSDIFF4_PER_CLK_PER_SMSP * sm_clk
        sad_avg = float(np.mean([me_sad_warp_instr(wl.tl(args.warmup + j)) for j in range(args.steps)]))
        me_avg_s = float(np.mean(stage_ms["me"])) * 1e-3
        me_issue = {"bound": "vabsdiff4_issue", "sad_warp_instr_per_launch": sad_avg, "peak_warp_instr_per_s": sad_peak,
                    "floor_ms_per_launch": sad_avg / sad_peak * 1e3, "avg_launch_ms": me_avg_s * 1e3, "frac": sad_avg / sad_peak / me_avg_s,
                    "source": "tools/ubench_pipes.cu: 0.49 VABSDIFF4 / clk / SM sub-partition (profiles/r02_ubench_pipes.txt)"}
        h2d = W * H * 3 // 2 * bps
        d2h = int(np.mean(d2h_bytes))
        la = np.array([[float(x) for x in t.tolist()] for t in link_all])
        kernel_names = {"me": "me_picture_kernel", "ois": "ois_picture_kernel", "mc": "mc_batch_kernel<1>" if hbd else "mc_batch8_kernel", "tq": "tq_batch_kernel",
                        "dlf": "dlf_pass_kernel", "h3_copy": "cudaMemcpy2DAsync (device to device)", "sao_stats": "sao_stats16_kernel" if hbd else "sao_stats2_kernel", "sao_dec": "sao_offsets_kernel + sao_merge_kernel",
                        "sao_apply": "sao_apply16_kernel" if hbd else "sao_apply_kernel", "pack": "pack_kernel", "pack_coeffs": "tq_pack_scan_kernel + tq_pack_copy_kernel"}
        out = {
            "metric": METRIC, "value": ctus / (ms * 1e-3), "unit": "CTU/s", "fps": ctus / (ms * 1e-3) / N_LCU,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16" if hbd else "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "stages": STAGES,
                       "l2": f"{RING}-frame ring of resident {W}x{H} pictures (> {RING * W * H * (1.33 + 1.5 * bps) / 1e6:.0f} MB of planes) > 126 MB L2",
                       "parallelism": f"{args.streams} compute stream(s) per GPU, {world} GPU(s)"},
            "e2e": {"value": ctus / (e2e_ms * 1e-3), "unit": "CTU/s", "fps": ctus / (e2e_ms * 1e-3) / N_LCU,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "host_link": pcie,
                    "host_link_all_ranks": {"h2d_gbs_min_max": [float(la[:, 0].min()), float(la[:, 0].max())], "d2h_gbs_min_max": [float(la[:, 1].min()), float(la[:, 1].max())],
                                            "duplex_each_gbs_min_max": [float(la[:, 2].min()), float(la[:, 2].max())]},
                    "numa_binding_rank0": numa, "kernel_launch": graph_note, "frames_in_flight": NS},
            "gpu_launches": int(launches),
            **({"diagnosis_only_e2e_skip": sorted(skip)} if skip else {}),
            "clocks": sampler.summary(),
            "roofline": {"kernel": kernel_names[dom],
                         "bound": "hbm", "achieved": alg[dom] / k_avg_s / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg[dom] / k_avg_s / 1e9 / peak, "traffic": traffic.get(dom), "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg[dom], "avg_launch_ms": k_avg_s * 1e3,
                         # achieved / frac above follow the contract: launch durations taken INSIDE the timed region, where the frames of several streams share the
                         # SMs (a launch lasts longer than the kernel's own time, the more streams the longer).  The same kernel alone on the GPU:
                         "alone": {"avg_launch_ms": float(np.mean(serial_ms[dom])), "achieved": alg[dom] / (float(np.mean(serial_ms[dom])) * 1e-3) / 1e9,
                                   "frac": alg[dom] / (float(np.mean(serial_ms[dom])) * 1e-3) / 1e9 / peak, "streams_in_timed_region": args.streams},
                         "note": "achieved / frac use launch durations inside the timed region, where the frames of several compute streams share the SMs: a launch lasts longer "
                                 "than the kernel's own time (round 1 ran 2 streams, this run --streams above: more streams raise `value` and lower this fraction); "
                                 "`alone` is the same kernel with nothing else on the GPU"},
            # SURVEY 8d's secondary bound for ME: the VABSDIFF4 instructions the search cannot avoid at the measured issue rate of their (half-rate) pipe
            "me_issue_roofline": me_issue,
            "timed_region_runs_ms": runs_ms,
            "stages": per_stage,
            "me_ms_by_temporal_layer": {str(tl): float(np.mean([t for j, t in enumerate(serial_ms["me"]) if wl.tl(args.warmup + j) == tl])) for tl in sorted(set(TL_PATTERN))
                                        if any(wl.tl(args.warmup + j) == tl for j in range(len(serial_ms["me"])))},
            "stages_serial": {n: {"launch_ms_avg": float(np.mean(v)) if v else None, "frac_of_hbm_peak": (alg[n] / (np.mean(v) * 1e-3) / 1e9 / peak) if v else None}
                              for n, v in serial_ms.items()},
        }
        if CFG_ID == 5:
            out["density"] = {"streams_per_gpu": args.streams, "streams_total": args.streams * world, "fps_per_stream": out["fps"] / (args.streams * world),
                              "realtime_streams_at_60fps": out["fps"] / 60.0}
        if ex is not None:
            xs = [a.elapsed_time(b) for a, b in ex_stats["ms"]]
            recv_bytes = ex_stats["bytes"]
            if ex.peer and ex_report:      # producers' view: every round the slowest producer's copies bound the round
                per = ex_report["per_rank"][1:]
                xs = [max(p["ms"][r] for p in per) for r in range(len(per[0]["ms"]))] if per and per[0]["ms"] else xs
            out["gop_exchange"] = {"frames_per_gop": FPG, "rounds": ex_stats["rounds"], "gops_total": ex_stats["rounds"] * world,
                                   "bytes_received_by_rank0": int(recv_bytes), "bytes_per_gop_avg": (recv_bytes / max(1, ex_stats["rounds"] * max(1, world - 1))),
                                   "transfer_ms_avg_on_comm_stream": float(np.mean(xs)) if xs else None,
                                   "nvlink_gbs_during_transfers": (recv_bytes / 1e9) / (sum(xs) * 1e-3) if (recv_bytes and xs and sum(xs) > 0) else None,
                                   "rank0_reassembled_coding_order": [g for g, _, _ in ex.assembled] == list(range(len(ex.assembled))),
                                   "transport": "peer memory (CUDA IPC mapping of rank 0's buffers, device-to-device copies over NVLink by the copy engines; torch.distributed carries the byte counts)"
                                                if ex.peer else "torch.distributed point-to-point send / recv",
                                   "payload_verified_on_rank0": ex_report["payload_verified"] if ex_report else None,
                                   "inside_timed_region": True,
                                   "link_gbs_lower_bound": (recv_bytes / 1e9) / (sum(runs_ms) * 1e-3) if recv_bytes else None,
                                   "note": "transfers are posted on a communication stream at every GOP boundary, overlap the next GOP's kernels and complete before the "
                                           "timed region ends; nvlink_gbs_during_transfers = bytes rank 0 received / time between posting and completion on rank 0's "
                                           "communication stream (includes waiting for the slowest sender's GOP); link_gbs_lower_bound = the same bytes / all timed time"}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_reference(frames=None)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own functions (oracle/_ref), one frame per host thread
# ------------------------------------------------------------------------------------------------
def host_info():
    """What the CPU arm really has: affinity mask, cgroup CPU quota, load, CPU model -- so that a record can explain its own CPU number"""
    info = {"os_cpu_count": os.cpu_count(), "affinity_threads": len(os.sched_getaffinity(0)), "loadavg_before": list(os.getloadavg())}
    quota = None
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            q, per = f.read().split()[:2]
            info["cgroup_cpu_max"] = f"{q} {per}"
            if q != "max":
                quota = max(1, int(float(q) / float(per)))
    except Exception:
        info["cgroup_cpu_max"] = None
    try:
        with open("/proc/cpuinfo") as f:
            info["cpu_model"] = next((l.split(":", 1)[1].strip() for l in f if l.startswith("model name")), None)
    except Exception:
        info["cpu_model"] = None
    info["usable_threads"] = max(1, min(info["affinity_threads"], quota or info["affinity_threads"]))
    return info


def cpu_reference(frames):
    """The reference's own functions on every host thread.  Like the B200 arm, the timed region holds nothing but the stage calls: the
    mode-decision stand-in (MC job lists from one ME pass per ring frame) is computed beforehand, every worker thread owns its output buffers for
    the whole run (no per-frame allocation), and the workers start together behind a barrier.  ctypes releases the GIL inside the calls."""
    import threading
    import common
    import svtb200 as S
    from concurrent.futures import ThreadPoolExecutor
    variant = "avx512" if common.cpu_has("avx512bw") else "avx2"
    ref = common.load_ref(variant)
    if ref is None:
        raise RuntimeError("oracle/_ref missing: build it with oracle/build_ref.sh where /root/reference exists")
    refsao = C.CDLL(os.path.join(ROOT, "oracle", "_ref", f"libsvtref_sao_{variant}.so"))
    # the threads this process may actually run on (affinity mask, cgroup CPU quota), not the box's nominal core count
    host = host_info()
    nthreads = host["usable_threads"]
    per_thread = 4 if W <= 4096 else 1
    frames = frames or nthreads * per_thread
    nworkers = min(nthreads, frames)
    hbd = BITS == 10
    sfx = "16" if hbd else ""
    wl = Workload(RING, seed=0x5F7C0001)
    pa = [S.HostPaPicture(l) for l in wl.luma]
    if hbd:
        yuv = [S.HostYuvPicture(wl.luma16(i), *wl.chroma16(i), dtype=np.uint16) for i in range(RING)]
    else:
        yuv = [S.HostYuvPicture(wl.luma[i], wl.chroma[i][0], wl.chroma[i][1]) for i in range(RING)]
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    ois_cfgs = {tl: wl.ois_cfg(tl) for tl in set(TL_PATTERN)}
    cap_ois = wl.ois_cap()
    f_mc, f_tq = getattr(ref, "svtref_mc_batch" + sfx), getattr(ref, "svtref_tq_batch" + sfx)
    f_dlfsao = getattr(ref, "svtref_dlf_sao_stats_picture" + sfx)
    f_stats, f_apply = getattr(refsao, "svtref_sao_stats_picture" + sfx), getattr(refsao, "svtref_sao_apply_picture" + sfx)
    refdec = C.CDLL(os.path.join(ROOT, "oracle", "_ref", f"libsvtref_saodec_{variant}.so"))
    f_dec = refdec.svtref_sao_decision_picture      # the reference's own DetermineSao*ModeOffsets / TestSaoCopyModes sequence (oracle/refinc_saodec_harness.c)
    rates = sao_rates()
    for f in (ref.svtref_me_picture_rows, ref.svtref_ois_picture, f_mc, f_tq, f_dlfsao, f_stats, f_apply, f_dec):
        f.restype = C.c_int
    flags = (C.c_uint8 * 2)(1, 1)
    me_cfgs = {tl: wl.me_cfg_tl(tl) for tl in set(TL_PATTERN)}
    ref_pairs = {}
    for i in range(RING):
        ref_pairs[wl.refs(i)] = S.yuv_array([yuv[wl.refs(i)[0]], yuv[wl.refs(i)[1]]])
    n_tq = len(wl.tq_jobs)
    sample_dt = np.uint16 if hbd else np.uint8

    def me(i, res):
        k = i % RING
        r0, r1 = wl.refs(i)
        assert ref.svtref_me_picture_rows(C.byref(me_cfgs[wl.tl(i)]), C.byref(pa[k].struct), C.byref(pa[r0].struct), C.byref(pa[r1].struct), vp(res), None, None, None,
                                          C.c_int(1), C.c_int(0), C.c_int(ROWS)) == 0

    # mode-decision stand-in, outside the timed region (the B200 arm does the same with its own ME): the job list of ring frame k
    def jobs_of(k):
        res = np.zeros((N_LCU, 85), dtype=S.ME_PU_RESULT_DTYPE)
        me(k, res)
        return wl.mc_jobs(res)
    with ThreadPoolExecutor(min(nthreads, RING)) as ex:
        mc_jobs = list(ex.map(jobs_of, range(RING)))

    class Worker:
        def __init__(self):
            self.res = np.zeros((N_LCU, 85), dtype=S.ME_PU_RESULT_DTYPE)
            self.ois_total, self.ois_cand = np.zeros((N_LCU, 85), np.uint8), np.zeros((N_LCU, 85, cap_ois), np.uint32)
            self.pred = S.HostYuvPicture(np.zeros((H, W), sample_dt), np.zeros((H // 2, W // 2), sample_dt), np.zeros((H // 2, W // 2), sample_dt), replicate=False,
                                         dtype=sample_dt)
            self.coeff = S.HostCoeffPicture(W, H)
            self.nz = np.zeros(n_tq, np.uint32)
            self.stats = np.zeros((N_LCU, 3), dtype=S.SAO_STATS_DTYPE)
            self.dec = np.zeros(N_LCU, dtype=S.SAO_DECISION_DTYPE)
            self.prm = np.zeros(N_LCU, dtype=S.SAO_PARAMS_DTYPE)

        def frame(self, i):
            k = i % RING
            tl = wl.tl(i)
            me(i, self.res)
            assert ref.svtref_ois_picture(C.byref(ois_cfgs[tl]), C.byref(pa[k].struct.full), vp(self.res), vp(self.ois_total), vp(self.ois_cand), C.c_uint32(cap_ois), C.c_int(1)) == 0
            jobs = mc_jobs[k]
            assert f_mc(ref_pairs[wl.refs(i)], C.c_int(2), vp(jobs), C.c_int(len(jobs)), C.byref(self.pred.struct), C.c_int(1)) == 0
            assert f_tq(C.byref(yuv[k].struct), C.byref(self.pred.struct), C.byref(self.pred.struct), C.byref(self.coeff.struct), vp(wl.tq_jobs),
                        C.c_int(n_tq), vp(self.nz), C.c_int(1), C.c_int(0), C.c_int(n_tq)) == 0
            if tl == 0:
                # the encoder's own order: every LCU's deblocking step followed by that LCU's SAO statistics (EbCodingLoop.c:4604-4635, 4711-4748)
                assert f_dlfsao(C.byref(self.pred.struct), C.byref(yuv[k].struct), vp(wl.vbs), vp(wl.hbs), vp(wl.qp_array), C.byref(wl.dlf_prm),
                                vp(self.stats), None, C.c_int(1)) == 0
                self.decide()
                self.prm["type"], self.prm["band_position"], self.prm["offset"] = self.dec["type"], self.dec["band_position"], self.dec["offset"]
                assert f_apply(C.byref(self.pred.struct), vp(self.prm), flags, C.c_int(1)) == 0
            else:
                assert f_stats(C.byref(yuv[k].struct), C.byref(self.pred.struct), vp(self.stats), C.c_int(1)) == 0
                self.decide()

        def decide(self):
            assert f_dec(vp(self.stats), C.c_int(COLS), C.c_int(ROWS), C.byref(rates), C.c_uint64(SAO_LAMBDA[0]), C.c_uint64(SAO_LAMBDA[1]), C.c_int(1 if hbd else 0), vp(self.dec)) == 0

    gate = threading.Barrier(nworkers + 1)
    errors = []

    def run(t):
        try:
            w = Worker()
            w.frame(t)                                  # warm-up: page in the libraries, touch this worker's buffers
            gate.wait()                                 # everybody is ready
            gate.wait()                                 # timed region starts
            for i in range(t, frames, nworkers):
                w.frame(i)
        except Exception as e:      # noqa: BLE001
            errors.append(e)
            gate.abort()
        finally:
            try:
                gate.wait()                             # timed region ends when the last worker is through
            except threading.BrokenBarrierError:
                pass

    threads = [threading.Thread(target=run, args=(t,)) for t in range(nworkers)]
    for th in threads:
        th.start()
    t0 = dt = None
    try:
        gate.wait()
        t0 = time.perf_counter()
        gate.wait()
        gate.wait()
        dt = time.perf_counter() - t0
    except threading.BrokenBarrierError:
        pass                                            # a worker failed and aborted the barrier: its exception is re-raised below
    for th in threads:
        th.join()
    if errors or dt is None:
        raise errors[0] if errors else RuntimeError("cpu_reference: the worker barrier broke")
    v = N_LCU * frames / dt
    return {"value": v, "unit": "CTU/s", "fps": v / N_LCU, "cores": nworkers, "kind": "reference", "host": host, "loadavg_after": list(os.getloadavg()),
            "sample": f"{frames} full {W}x{H} frames of the same stages and jobs through the reference's own functions (MotionEstimateLcu, OpenLoopIntraSearchLcu, "
                      f"EncodePassInterPrediction{sfx}, EstimateTransform/UnifiedQuantizeInvQuantize/EncodeInvTransform, the LCU DLF drivers interleaved with the SAO gather "
                      f"on layer-0 pictures, SAO gather/apply; {variant} rows), {nworkers} worker threads with their own buffers, {frames // nworkers} frame(s) each behind a "
                      f"barrier; job lists prepared beforehand as on the B200 arm"}


def reference_encoder_app(n_frames=32, timeout_s=180):
    """The reference's WHOLE encoder (its own CLI, AVX-512 build from oracle/build_ref.sh) on the same synthetic 4K video at preset M10, CQP:
    ME + mode decision + EncDec + entropy coding on all host threads.  Informational: the stages this repo replaces are timed by cpu_reference();
    this is what the unmodified encoder delivers end to end on the box.  Never raises (the line it decorates must not fail because of it)."""
    import re
    import subprocess
    import tempfile
    app = os.path.join(ROOT, "oracle", "_ref", "SvtHevcEncApp_avx512")
    try:
        import common
        if not os.path.exists(app) or not common.cpu_has("avx512bw"):
            return {"unavailable": "oracle/_ref/SvtHevcEncApp_avx512 missing or no AVX-512 on this host"}
        if BITS != 8:
            return {"unavailable": "whole-encoder context run is wired for the 8-bit configurations only"}
        wl = Workload(RING, seed=0x5F7C0001)
        with tempfile.TemporaryDirectory() as d:      # -nb preloads the frames into RAM, so the file system does not matter
            yuv = os.path.join(d, "in.yuv")
            with open(yuv, "wb") as f:
                for i in range(n_frames):
                    k = i % RING
                    f.write(wl.luma[k].tobytes()); f.write(wl.chroma[k][0].tobytes()); f.write(wl.chroma[k][1].tobytes())
            cmd = [app, "-i", yuv, "-w", str(W), "-h", str(H), "-n", str(n_frames), "-nb", str(n_frames), "-encMode", str(ENC_MODE), "-q", "32", "-fps", str(FPS),
                   "-intra-period", "55", "-irefresh-type", "1", "-b", os.path.join(d, "out.265")]
            t0 = time.perf_counter()
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s).stdout
            wall = time.perf_counter() - t0
        m = re.search(r"Average Speed:\s*([0-9.]+)\s*fps", out)
        if not m:
            return {"unavailable": "no 'Average Speed' line in the encoder's output", "tail": out[-300:]}
        fps = float(m.group(1))
        return {"fps": fps, "ctu_per_s": fps * N_LCU, "frames": n_frames, "threads": os.cpu_count(), "wall_s": wall,
                "what": f"SvtHevcEncApp (reference, AVX-512 build) {W}x{H} 8-bit, -encMode {ENC_MODE} -q 32 -intra-period 55, frames preloaded (-nb): whole encoder"}
    except Exception as e:      # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    t0 = time.perf_counter()
    cb = cpu_reference(frames=None)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "CTU/s", "fps": cb["fps"], "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * N_LCU / cb["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u16" if BITS == 10 else "u8", "data": "synthetic", "config": {"workload": WORKLOAD, "stages": STAGES},
           "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "CTU/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": time.perf_counter() - t0}
    if not args.no_encoder_app:
        out["reference_encoder_app"] = reference_encoder_app()
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-encoder-app", action="store_true", help="reference arm: skip the whole-encoder run of the reference's own CLI")
    ap.add_argument("--e2e-skip", default="", help="diagnosis only: comma list of e2e parts to leave out (h2d, d2h, pack, pa); a run with this set is not a bench value")
    ap.add_argument("--inflight", type=int, default=5, help="frames in flight in the end-to-end loop (device + pinned host buffers per frame)")
    ap.add_argument("--streams", type=int, default=3, help="compute streams the resident loop spreads consecutive frames over")
    ap.add_argument("--repeats", type=int, default=5, help="the timed region (--steps frames) is run this many times; the median run is reported")
    ap.add_argument("--no-graphs", action="store_true", help="e2e loop: launch every stage call by call instead of replaying one captured CUDA graph per frame")
    ap.add_argument("--graphs", action="store_true", help="e2e loop: replay captured frame graphs even where the configuration's default is call by call (measured per "
                    "configuration: the 1080p and the 10-bit 4K loops are host-bound and gain 6-18 %, the 8-bit 4K loop is not and loses 7 %)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "p2p"], help="config 4, N > 1: GOP hand-off by peer-memory copies over NVLink (default) or by send / recv")
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config (1-based index in `configs`): 2 headline (default), 3 10-bit M7 4K, "
                    "4 8K 10-bit closed GOPs over the GPUs with the NCCL GOP-boundary exchange, 5 1080p stream density")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    c = set_config(args.config)
    if "--streams" not in sys.argv:
        args.streams = c["streams"]
    if not args.graphs and not c["graphs"]:
        args.no_graphs = True
    if args.config == 4 and "--steps" not in sys.argv:
        args.steps = 32
    if args.config == 4 and "--repeats" not in sys.argv:
        args.repeats = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""bench.py -- BASELINE.json metric: 4K YUV420 encode fps @ fixed QP (-preset veryfast -rc 0 -qp 27 -iper 128).

  python bench.py --gpus N --steps K --warmup W [--config 4k]        # this repo's B200 hot path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...             # the reference's own CPU encoder (oracle/_ref/appencoder)
  --config 1080p | 4k (default, the metric's config) | 4k_slow_crf | 8k   = BASELINE.json configs[1..4]

A step = every stream of the rank encodes one GOP shard (1 IDR + iper-1 P, closed GOP) of synthetic I420; every stream has its OWN input
sequence (the base sequence shifted cyclically by a stream-specific offset), in HBM for `value` and in pinned host memory for `e2e`.
  value : device hot path only (ME + sub-pel, CU/merge decision, MC + DCT/quant/IDCT, deblock, SAO, level packing, syntax D2H) with the
          pictures already resident in HBM and read in place -- whole-job pictures/s over all ranks.
  e2e   : the same GOP shards through the public encoder API with HOST buffers: H2D of every picture, the device hot path, D2H of the
          frame syntax, host CABAC -> Annex-B bytes (and, for N>1, the NCCL gather of the NAL units).
Both arms report bitrate (kbps at 30 fps) and luma PSNR next to fps: the two encoders do not make the same decisions.
Timing: CUDA events after a device-wide synchronize on both sides (all streams idle), barrier before, max over ranks.
Working set per step (streams x iper x picture bytes, distinct per stream) is far larger than the 126 MB L2: no flush needed.
"""
import argparse
import json
import math
import os
import re
import shutil
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

CONFIGS = {
    "1080p": dict(w=1920, h=1080, preset="veryfast", rc=0, qp=27, crf=None, iper=128, baseline="BASELINE configs[1]"),
    "4k": dict(w=3840, h=2160, preset="veryfast", rc=0, qp=27, crf=None, iper=128, baseline="BASELINE configs[2] (the metric's configuration)"),
    "4k_slow_crf": dict(w=3840, h=2160, preset="slow", rc=3, qp=27, crf=24.0, iper=128, baseline="BASELINE configs[3]"),
    "8k": dict(w=7680, h=4320, preset="veryfast", rc=0, qp=27, crf=None, iper=32, baseline="BASELINE configs[4] (32-picture shards: 1.6 GB of input each)"),
}
DISTINCT = 16                      # distinct synthetic pictures; a shard plays them forward/backward (smooth motion)
FPS_NOMINAL = 30.0


def workload_name(c):
    rc = "-rc 0 -qp %d" % c["qp"] if c["rc"] == 0 else "-rc 3 -crf %g" % c["crf"]
    return "%dx%d I420 -preset %s %s -iper %d" % (c["w"], c["h"], c["preset"], rc, c["iper"])


def shard_order(n, distinct=DISTINCT):
    """ping-pong index sequence 0..d-1, d-2..0, 1.. of length n: consecutive pictures always differ by exactly one motion step"""
    seq, i, d = [], 0, 1
    for _ in range(n):
        seq.append(i)
        if i + d < 0 or i + d >= distinct:
            d = -d
        i += d
    return seq


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def bind_to_gpu_numa(gpu_index, min_cpus):
    """pin this process (and the shard threads it will start, and the pinned buffers it will first-touch) to the CPUs NVML reports as local to
    the GPU: kernel launches and H2D/D2H descriptors then stay on the GPU's own socket.  Returns (n_cpus, previous mask) or (0, None)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        prev = os.sched_getaffinity(0)
        cpus &= prev
        if len(cpus) >= min_cpus and cpus != prev:
            os.sched_setaffinity(0, cpus)
            return len(cpus), prev
    except Exception:
        pass
    return 0, None


def prefer_gpu_numa_memory(gpu_index):
    """memory policy only (no CPU pinning): page-locked buffers allocated from now on come from the NUMA node the GPU hangs off, so the
    H2D stream of each rank stays on its own socket's memory controller instead of crossing the inter-socket link (what `numactl --preferred`
    does).  Returns the node or None."""
    try:
        import ctypes
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(gpu_index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus[-12:]).read())
        if node < 0:
            return None
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        libc = ctypes.CDLL(None, use_errno=True)
        if libc.syscall(238, 1, mask, 16 * 64 + 1) != 0:          # set_mempolicy(MPOL_PREFERRED, mask, maxnode), x86-64 syscall 238
            return None
        return node
    except Exception:
        return None


def run_cli(binary, c, frames, yuv_path, extra=()):
    """run an AppEncoder-compatible CLI (the reference binary or ours); returns dict(fps, wall, kbps, psnr_y) from its own summary lines"""
    out = "/dev/shm/ks265_cli_%d.265" % os.getpid()
    rc = ["-rc", "0", "-qp", str(c["qp"])] if c["rc"] == 0 else ["-rc", "3", "-crf", str(c["crf"])]
    cmd = [binary, "-i", yuv_path, "-wdt", str(c["w"]), "-hgt", str(c["h"]), "-fr", str(int(FPS_NOMINAL)), "-preset", c["preset"], *rc,
           "-iper", str(c["iper"]), "-frms", str(frames), "-psnr", "1", "-b", out, *extra]
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=3600)
    wall = time.time() - t0
    if os.path.exists(out):
        os.remove(out)
    m = re.search(r"Total Frames:\s*(\d+),\s*test time:\s*([\d.]+)\s*ms,\s*FPS:\s*([\d.]+)", r.stdout)
    if not m:
        raise RuntimeError("%s gave no timing line: %s %s" % (os.path.basename(binary), r.stdout[-400:], r.stderr[-400:]))
    q = re.search(r"bitrate, psnr:\s*([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)", r.stdout)
    return {"fps": float(m.group(3)), "wall_s": wall, "kbps": float(q.group(1)) if q else None, "psnr_y": float(q.group(2)) if q else None}


def base_sequence(c):
    """the base GOP shard: list of iper uint8 pictures (views of DISTINCT generated ones)"""
    import numpy as np
    import gen_yuv
    frames = [np.frombuffer(fr, np.uint8) for fr in gen_yuv.frames(c["w"], c["h"], DISTINCT, seed=1234)]
    return [frames[i] for i in shard_order(c["iper"])]


def write_yuv(seq, n, path, repeat=1):
    with open(path, "wb") as f:
        for _ in range(repeat):
            for fr in seq[:n]:
                f.write(fr.tobytes())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="4k", choices=sorted(CONFIGS))
    ap.add_argument("--streams", type=int, default=0, help="concurrent GOP shards per GPU (0 = auto)")
    ap.add_argument("--cpu-sample-frames", type=int, default=0, help="pictures per reference run (0 = one whole GOP shard)")
    ap.add_argument("--no-cli", action="store_true", help="skip the CLI-binary end-to-end leg")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    W, H, IPER = c["w"], c["h"], c["iper"]
    FSZ = W * H * 3 // 2
    METRIC = "4K YUV420 encode fps @ fixed QP (3840x2160 -preset veryfast -rc 0 -qp 27 -iper 128)" if a.config == "4k" else "encode fps, " + workload_name(c)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "appencoder")
    ref_frames = a.cpu_sample_frames or IPER

    import numpy as np
    t_start = time.time()

    def note(msg):          # progress marks on stderr (the JSON line on stdout stays alone)
        print("[bench %6.1fs] %s" % (time.time() - t_start, msg), file=sys.stderr, flush=True)

    if a.impl == "reference":
        if rank != 0:
            return 0
        seq = base_sequence(c)
        path = "/dev/shm/ks265_bench_%d.yuv" % os.getpid()
        write_yuv(seq, ref_frames, path)
        variants = {}
        try:
            runs = []
            for i in range(a.warmup + a.steps):
                r = run_cli(ref_bin, c, ref_frames, path, ("-threads", "0"))
                if i >= a.warmup:
                    runs.append(r)
            # the same clip at matched picture structure (our streams are IDR + P...), and the AVX2 build of the reference
            try:
                variants["centos_x64 -bframes 0 (P-only, the structure this repo emits)"] = run_cli(ref_bin, c, ref_frames, path, ("-threads", "0", "-bframes", "0"))
                avx2 = os.path.join(ROOT, "oracle", "_ref", "appencoder_avx2")
                if os.path.exists(avx2):
                    variants["ubuntu_x64 (AVX2 build), default GOP"] = run_cli(avx2, c, ref_frames, path, ("-threads", "0"))
                variants["centos_x64 -threads 1"] = run_cli(ref_bin, c, min(ref_frames, 32), path, ("-threads", "1"))
            except Exception as ex:
                variants["error"] = str(ex)[:200]
        finally:
            os.remove(path)
        fps = sum(r["fps"] for r in runs) / len(runs)
        ms = 1000.0 * sum(r["wall_s"] for r in runs) / len(runs)
        sample = "%d pictures (one GOP shard) of the same synthetic sequence per step, centos_x64/appencoder -threads 0 (all %d host cores), stock %s settings " \
                 "(hierarchical-B GOP, lookahead), fps from its own 'test time' line, %d runs: %s" % (ref_frames, cores, c["preset"], len(runs), ["%.1f" % r["fps"] for r in runs])
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": {"workload": workload_name(c), "baseline_config": c["baseline"], "sample_frames": ref_frames},
                          "quality": {"kbps": runs[-1]["kbps"], "psnr_y": runs[-1]["psnr_y"], "fps_nominal": FPS_NOMINAL},
                          "reference_variants": variants,
                          "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": sample},
                          "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import torch.distributed as dist
    import ks265codec_b200 as ks
    from ks265codec_b200 import shard as ksh
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # shards in flight per GPU, measured on the 64-core/128-thread pool host at N=1 (value / e2e fps): 8 -> e2e host-bound, 16 -> 2625 / 2487,
    # 24 with sleeping waits -> 2614 / 2601 (with spinning waits 24 shards collapse to 2037 / 1813: the spinners starve the launch threads)
    # shards in flight per GPU: the device saturates at ~12-16 (measured: 16 / 24 / 32 give the same fps at N=1) and a shard thread sleeps while its
    # picture is on the device and needs ~1.3 ms of host time per 4K picture, so up to 4 threads per core are fine when GPUs outnumber cores / 16
    cap = 24 if W * H <= 3840 * 2160 else 12
    per_gpu = cores // max(1, a.gpus)
    streams = a.streams or max(2, min(cap, max(per_gpu, min(16, 4 * per_gpu))))
    visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    nvml_index = int(visible.split(",")[local_rank]) if visible and all(v.strip().isdigit() for v in visible.split(",")) else local_rank
    # opt-in: measured neutral on the 2-socket pool hosts (N=2: value 4403 bound vs 4414 unbound, e2e 3638 vs 3838), profiles/README.md
    numa_cpus, full_mask = bind_to_gpu_numa(nvml_index, 2 * streams) if os.environ.get("KS_NUMA_BIND") == "1" else (0, None)
    numa_node = prefer_gpu_numa_memory(nvml_index) if os.environ.get("KS_NUMA_MEM", "1") == "1" else None
    # shard threads SLEEP while they wait for a picture (cudaEventBlockingSync).  Spinning waits are 2 % faster only for <= 16 shards of ONE
    # process; beyond that, and whenever several GPU processes share the host, the spinners slow everybody's launch threads
    # (N=2, 16 shards per GPU: 4414 / 3838 fps spinning vs 5226 / 4440 sleeping).
    os.environ.setdefault("KS_BLOCKING_SYNC", "1")

    # ---- synthetic input: one sequence PER STREAM = the base shard shifted cyclically by a stream-specific offset (different content at
    #      every CTU position, different addresses); device copies for `value`, pinned host copies for `e2e` ----
    note("generating the synthetic sequence")
    seq = base_sequence(c)
    note("staging %d per-stream sequences (HBM + pinned host)" % streams)
    base = torch.empty(IPER * FSZ, dtype=torch.uint8, pin_memory=True)
    bv = base.numpy()
    for k, fr in enumerate(seq):
        bv[k * FSZ:(k + 1) * FSZ] = fr
    base_dev = base.cuda()

    def shifted(s):
        if s == 0:
            return base_dev
        t = base_dev.view(IPER, FSZ)
        dy, dx = 2 * ((37 * (s + rank * streams)) % (H // 2)), 2 * ((101 * (s + rank * streams)) % (W // 2))
        y = torch.roll(t[:, :W * H].view(IPER, H, W), (dy, dx), (1, 2)).reshape(IPER, -1)
        u = torch.roll(t[:, W * H:W * H * 5 // 4].view(IPER, H // 2, W // 2), (dy // 2, dx // 2), (1, 2)).reshape(IPER, -1)
        v = torch.roll(t[:, W * H * 5 // 4:].view(IPER, H // 2, W // 2), (dy // 2, dx // 2), (1, 2)).reshape(IPER, -1)
        return torch.cat([y, u, v], 1).reshape(-1).contiguous()
    dev_seqs = [shifted(s) for s in range(streams)]
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    distinct_host = avail > 3 * streams * IPER * FSZ * max(1, world)
    if distinct_host:
        host_seqs = [base] + [torch.empty(IPER * FSZ, dtype=torch.uint8, pin_memory=True) for _ in range(1, streams)]
        for s in range(1, streams):
            host_seqs[s].copy_(dev_seqs[s])
    else:
        host_seqs = [base] * streams
    host_np = [t.numpy() for t in host_seqs]
    torch.cuda.synchronize()

    note("opening %d encoders" % streams)
    kw = dict(preset=c["preset"], qp=c["qp"], iper=IPER, device=local_rank, psnr=1, rc=c["rc"])
    if c["crf"] is not None:
        kw["crf"] = c["crf"]
    cfg = ks.default_config(W, H, **kw)
    encs = [ks.Encoder(cfg) for _ in range(streams)]
    outs = [np.empty(IPER * FSZ // 4 + (8 << 20), np.uint8) for _ in range(streams)]      # Annex-B output buffers
    results = [None] * streams

    def step_device():
        def work(i):
            results[i] = encs[i].run_gop_device(dev_seqs[i].data_ptr(), IPER)
        th = [threading.Thread(target=work, args=(i,)) for i in range(streams)]
        [t.start() for t in th]; [t.join() for t in th]

    def step_e2e():
        def work(i):
            results[i] = encs[i].encode_gop(host_np[i], want_recon=False, nframes=IPER, out=outs[i])
        th = [threading.Thread(target=work, args=(i,)) for i in range(streams)]
        [t.start() for t in th]; [t.join() for t in th]
        if world > 1:       # the only exchange of the job: NAL units of every shard to rank 0 over NCCL
            local = {rank + world * i: results[i][0] for i in range(streams)}
            ksh.gather_bitstreams(local, world * streams, as_array=True)

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1)
        clk = sampler.stop() if sampler else None
        if world > 1:
            t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms, clk

    # ---- value: device hot path, inputs resident in HBM ----
    note("value arm")
    for e in encs:
        e.set_profiling(True)
    ms_dev, clocks = timed(step_device, a.steps, a.warmup, ClockSampler(local_rank) if rank == 0 else None)
    launches = sum(int(r.gpu_launches) for r in results) * a.steps      # every step issues the same launches
    stage = {}
    for e in encs:
        for k, (ms, n) in e.stage_times().items():
            t = stage.setdefault(k, [0.0, 0]); t[0] += ms; t[1] += n
        e.set_profiling(False)
    frames_per_step = streams * IPER * world
    value = frames_per_step * a.steps / (ms_dev / 1000.0)

    # ---- e2e: host buffers -> Annex-B bytes through the public API ----
    note("e2e arm")
    ms_e2e, _ = timed(step_e2e, a.steps, max(1, min(a.warmup, 1)))
    e2e = frames_per_step * a.steps / (ms_e2e / 1000.0)
    h2d = sum(int(r[2].h2d_bytes) for r in results)
    d2h = sum(int(r[2].d2h_bytes) for r in results)
    bs_bytes = sum(int(r[2].bytes) for r in results)
    sse_y = sum(int(r[2].sse[0]) for r in results)
    all_streams = {"kbps": bs_bytes * 8.0 * FPS_NOMINAL / (streams * IPER) / 1000.0,
                   "psnr_y": 10.0 * math.log10(255.0 ** 2 * W * H * streams * IPER / max(1, sse_y)),
                   "note": "this rank's %d shards of the e2e arm: shard 0 is the clip the reference arm encodes, the others are that clip rolled cyclically by a "
                           "shard-specific offset (distinct inputs), which wraps its moving objects through the picture border -- harder content than the clip itself" % streams}
    # like for like with the reference arm: shard 0 of every rank is the unshifted clip, the one `--impl reference` / cpu_baseline encode
    try:
        r0 = results[0][2]
        quality = {"kbps": int(r0.bytes) * 8.0 * FPS_NOMINAL / IPER / 1000.0,
                   "psnr_y": 10.0 * math.log10(255.0 ** 2 * W * H * IPER / max(1, int(r0.sse[0]))), "fps_nominal": FPS_NOMINAL,
                   "note": "shard 0 of the e2e arm = the same %d-picture clip the reference arm encodes; bitrate at a nominal %g fps; PSNR from the device's SSE over the display area" % (IPER, FPS_NOMINAL),
                   "all_streams": all_streams}
    except Exception as ex:      # never let the quality annotation cost the bench line
        quality = dict(all_streams, fps_nominal=FPS_NOMINAL, error=str(ex)[:100])

    # ---- roofline of the dominant stage (SURVEY.md 8d per-kernel algorithmic bytes; S = 1.5*W*H per picture) ----
    S = 1.5 * W * H
    # algorithmic bytes per launch: compulsory HBM traffic with perfect on-chip reuse
    alg = {"me": S + S + S + 12.0 * W * H / 256,                # source luma+ (S_luma) + reference picture -> MV field + distortions + prediction planes
           "decide": 2.0 * W * H + 24.0 * W * H / 256,          # source luma + reference luma, search field in, final cells out (re-predicted cells extra)
           "recon_inter": S + S + S + 2.0 * S + 8.0 * W * H / 256,    # source + prediction in, reconstruction + int16 levels out
           "recon_intra": S + S + 2.0 * S, "intra_p": 16.0 * W * H / 256, "deblock": 2.0 * W * H, "sao": 3.0 * S, "pack": 2.0 * S + 0.1 * S}
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the 4K configuration (profiles/);
    # ncu flushes L2 before each replay, writes mostly stay in the 126 MB L2, so traffic can be BELOW the algorithmic bytes
    ncu_traffic = {"me": 21.9e6, "recon_inter": 32.0e6, "sao": 25.0e6, "deblock": 12.9e6, "pack": 31.6e6, "recon_intra": None, "decide": None} if a.config == "4k" else {}
    tr_path = os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")
    if a.config == "4k" and os.path.exists(tr_path):
        try:
            ncu_traffic.update(json.load(open(tr_path)))
        except Exception:
            pass
    # every stage timed ALONE (one stream, nothing else on the GPU): these are the launch durations the roofline uses.
    # dominant stage = largest solo time per picture, weighted by how often the stage runs in a GOP shard.
    solo = encs[0]
    solo.set_profiling(True)
    torch.cuda.synchronize()
    solo.run_gop_device(dev_seqs[0].data_ptr(), min(24, IPER))
    torch.cuda.synchronize()
    solo_t = solo.stage_times()
    solo.set_profiling(False)
    per_pic = {k: (v[0] / v[1] if v[1] else 0.0) * ((1.0 / IPER) if k == "recon_intra" else ((IPER - 1.0) / IPER if k in ("me", "decide", "recon_inter", "intra_p") else 1.0))
               for k, v in solo_t.items()}
    dom = max(per_pic, key=per_pic.get)
    avg_ms = solo_t[dom][0] / max(1, solo_t[dom][1])
    peak, peak_src = peaks()
    achieved = alg[dom] / (avg_ms / 1000.0) / 1e9
    pipeline_alg = S * 8.02                                 # B_alg of a P picture = S * (7.02 + R), R = 1
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic.get(dom),
                "peak_source": peak_src, "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": alg[dom],
                "pipeline": {"algorithmic_bytes_per_picture": pipeline_alg, "achieved": pipeline_alg * value / max(1, world) / 1e9, "frac": pipeline_alg * value / max(1, world) / 1e9 / peak},
                "stage_ms_share": {k: v[0] / max(1e-9, sum(x[0] for x in stage.values())) for k, v in stage.items()},
                "solo_stage_ms": {k: (v[0] / v[1] if v[1] else None) for k, v in solo_t.items()},
                "solo_ms_per_picture": per_pic,
                "per_stage": {k: {"achieved": alg[k] / (solo_t[k][0] / solo_t[k][1] / 1000.0) / 1e9, "frac": alg[k] / (solo_t[k][0] / solo_t[k][1] / 1000.0) / 1e9 / peak}
                              for k in solo_t if solo_t[k][1] and solo_t[k][0] > 0},
                "limiter": {"me": "issue slots (integer SAD/interpolation ALU), not HBM", "decide": "issue slots (interpolation of the candidate vectors) + the serial quadtree decision, not HBM",
                            "recon_inter": "issue slots + barriers (integer transforms), not HBM",
                            "recon_intra": "dependency chain of the CTU wavefront", "intra_p": "dependency chain / latency (few cells)", "sao": "shared-memory/ALU, then HBM", "deblock": "latency", "pack": "HBM"}[dom],
                "note": "dominant stage = largest solo time per picture (GOP-weighted); avg_launch_ms = that stage timed alone on an idle GPU with CUDA events on its stream (%d pictures, one stream); peak = burst copy bandwidth. stage_ms_share = event intervals inside the timed region with %d streams sharing the GPU (includes co-scheduling waits)." % (min(24, IPER), streams)}

    for e in encs:
        e.close()
    del encs, dev_seqs
    torch.cuda.empty_cache()

    # ---- cpu baseline + the CLI-binary end-to-end leg (rank 0, N=1 only) ----
    note("cpu baseline / CLI leg")
    cpu, cli = None, None
    if rank == 0 and world == 1:
        path = "/dev/shm/ks265_bench_%d.yuv" % os.getpid()
        try:
            write_yuv(seq, ref_frames, path)
            if full_mask:
                os.sched_setaffinity(0, full_mask)      # the reference gets every host core, not just this GPU's socket
            r = run_cli(ref_bin, c, ref_frames, path, ("-threads", "0"))
            cpu = {"value": r["fps"], "unit": "frames/s", "cores": cores, "kind": "reference", "kbps": r["kbps"], "psnr_y": r["psnr_y"],
                   "sample": "%d pictures (one GOP shard) of the same synthetic sequence, oracle/_ref/appencoder (centos_x64) -threads 0, stock %s settings, fps from its 'test time' line (wall %.1f s)" % (ref_frames, c["preset"], r["wall_s"])}
        except Exception as ex:          # the baseline is reported, never required for the GPU numbers
            cpu = {"value": None, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": "failed: %s" % str(ex)[:200]}
        if not a.no_cli:
            try:
                # the drop-in boundary itself: our appencoder binary reading a YUV file (pageable memory), 8 GOP shards so its shard pipeline fills
                free = shutil.disk_usage("/dev/shm").free
                rep = max(1, min(8, int(free * 0.6 // (IPER * FSZ))))
                write_yuv(seq, IPER, path, repeat=rep)
                r = run_cli(ks.CLI_PATH, c, IPER * rep, path, ("-streams", str(min(streams, rep))))
                cli = {"value": r["fps"], "unit": "frames/s", "kbps": r["kbps"], "psnr_y": r["psnr_y"],
                       "scope": "ks265codec_b200/bin/appencoder (the AppEncoder drop-in) on a %d-picture YUV file in /dev/shm: file read into pageable memory, H2D, device, CABAC, Annex-B file written; %d shards in flight; fps from its 'test time' line (includes encoder open/close)" % (IPER * rep, min(streams, rep))}
            except Exception as ex:
                cli = {"value": None, "scope": "failed: %s" % str(ex)[:200]}
        if os.path.exists(path):
            os.remove(path)
    note("done")
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(c) + " (%s)" % c["baseline"], "gop_shard": "1 IDR + %d P, closed GOP" % (IPER - 1),
                       "streams_per_gpu": streams, "cpu_binding": ("%d CPUs local to the GPU (NVML affinity)" % numa_cpus) if numa_cpus else "none", "host_memory_node": numa_node, "pictures_per_step": frames_per_step, "parallelism": "gop-shard x%d" % world,
                       "l2": "inputs larger than L2: %d streams x %d distinct sequences of %.2f GB each per GPU (%s host buffers), read in place" % (streams, streams, IPER * FSZ / 1e9, "distinct pinned" if distinct_host else "one shared pinned"),
                       "value_scope": "device hot path (ME, CU/merge decision, MC+transform+quant, deblock, SAO, level pack, syntax D2H), pictures resident in HBM; host CABAC excluded",
                       "e2e_scope": "host I420 -> H2D -> device hot path -> syntax D2H -> host CABAC -> Annex-B (+NCCL NAL gather if N>1)",
                       "same_decisions_as_reference": "no: this repo emits IDR + P (4-picture QP cascade), the reference arm runs its stock preset (hierarchical-B GOP, lookahead); `quality` and cpu_baseline.kbps/psnr_y put the two rate-distortion points side by side"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e / a.steps,
                    "bitstream_bytes_per_step": bs_bytes * world},
            "e2e_cli": cli, "quality": quality,
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

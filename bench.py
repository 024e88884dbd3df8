#!/usr/bin/env python
"""bench.py -- encode/decode GSamples/s of the FLAC hot path on B200 (driver contract).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c3|c5|c4]

Workload (default c3 = BASELINE.json configs[2], the configuration the encode-scaling metric is
quoted on; it fits one GPU): synthetic Sentinel-2-like 10980x10980 uint16, 8 bands, streaming
tile_size 1024 -> 121 tiles / 964 483 200 samples, compression level 5.  A "step" is one pass of
the encode hot path over that raster, resident in HBM: per-tile min/max -> normalise -> subframe
analysis + Rice coding -> frame assembly (+ the size read-back that fixes the byte offsets).
After the K timed encode steps the decode direction (sync scan -> Rice decode/LPC restore ->
CRC-16 -> denormalise) is timed the same way and reported under "decode".
N > 1: one process per GPU, each rank encodes its own 121-tile scene (weak scaling: the job is N
scenes' tiles sharded by tile), the only collective is the all-gather of per-tile sizes.
`e2e` is the same step through the public engine API with HOST (pinned) buffers: H2D of the
raster and D2H of the frames inside the timed region.
`--impl reference` times the reference's CPU implementation of the path (the oracle port of
libFLAC's procedure + the reference's numpy normalisation, one tile per task over all host
threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "encode GSamples/s (streaming tiles -> FLAC frames; decode GSamples/s reported beside it)"
UNIT = "GSamples/s"

WORKLOADS = {
    # name: (description, level, tile_size, bands)
    "c3": ("synthetic Sentinel-2-like 10980x10980 uint16 x 8 bands, streaming tile_size 1024 (121 tiles), level 5", 5, 1024, 8),
    "c5": ("4096 tiles of 512x512 int16 (batch decode sweep corpus), level 5", 5, 512, 1),
    "c4": ("synthetic float32 DEM 32768x32768 via normalize_to_audio, level 8, single stream", 8, 32768, 1),
}


def make_raster(workload: str, device, scale_div: int = 1):
    from flac_raster_b200 import synth
    if workload == "c3":
        side = 10980 // scale_div
        return synth.sentinel2_like(side, side, 8, device=device)
    if workload == "c5":
        return synth.dem_int16_tiles(4096 // (scale_div * scale_div), 512, device=device)
    side = 32768 // scale_div
    return synth.dem_float32(side, side, device=device)


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region (nvidia-smi, 200 ms)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        # NVML in-process (a query takes well under a millisecond, so even a 50 ms timed region gets several samples);
        # nvidia-smi subprocesses (~0.2 s per sample) only if the binding is missing
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown if hasattr(pynvml, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                    "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._halt.is_set():
                try:
                    self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    r = int(get_reasons(h))
                    for n, b in bits.items():
                        if r & int(b):
                            self.reasons.add(n)
                except Exception:  # noqa: BLE001
                    pass
                self._halt.wait(0.004)
            return
        except Exception:  # noqa: BLE001
            pass
        import subprocess
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().splitlines()
                if out:
                    f = [x.strip() for x in out[0].split(",")]
                    self.samples.append(float(f[0]))
                    self.max_mhz = float(f[1])
                    for n, v in zip(names, f[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.2)

    def _nvml_index(self) -> int:
        # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it lists plain indices
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            ids = [int(v) for v in vis.split(",") if v.strip() != ""]
            if ids and self.index < len(ids):
                return ids[self.index]
        except ValueError:
            pass
        return self.index

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_tile_encode(args):
    """The reference's per-tile path (cli.py:558-598 -> converter.py:99-154) on CPU:
    interleave -> normalize_to_audio (numpy) -> libFLAC-procedure encode (oracle C port)."""
    win, level = args
    from oracle import flac_oracle as fo, normalization_oracle as no
    bands, h, w = win.shape
    inter = win.transpose(1, 2, 0).reshape(-1, bands)
    rate, bits = no.calculate_audio_params((h, w), win.dtype)
    audio, _ = no.normalize_to_audio(inter, bits)
    enc, fs = fo.encode(audio, 16 if bits == 16 else 32, rate, level)
    return inter.size, len(enc), enc


def cpu_tile_decode(args):
    enc, dtype, dmin, dmax, scale = args
    from oracle import flac_oracle as fo, normalization_oracle as no
    pcm, _ = fo.decode(enc)
    a = pcm.astype(np.int16) if scale == 32767 else pcm
    out = no.denormalize_from_audio(a, dmin, dmax, dtype, scale)
    return pcm.size


def sample_windows(workload: str, n_tiles: int, device):
    """First n_tiles tiles of the workload as host arrays (same content as the full raster)."""
    from flac_raster_b200 import synth
    desc, level, tile, bands = WORKLOADS[workload]
    wins = []
    if workload == "c3":
        per_row = (10980 + tile - 1) // tile
        rows = (n_tiles + per_row - 1) // per_row
        r = synth.sentinel2_like(min(10980, rows * tile), 10980, 8, device=device).cpu().numpy()
        for i in range(n_tiles):
            ty, tx = divmod(i, per_row)
            wins.append(np.ascontiguousarray(r[:, ty * tile:(ty + 1) * tile, tx * tile:min(10980, (tx + 1) * tile)]))
    elif workload == "c5":
        r = synth.dem_int16_tiles(n_tiles, 512, device=device).cpu().numpy()
        wins = [np.ascontiguousarray(r[:, i * 512:(i + 1) * 512]) for i in range(n_tiles)]
    else:
        side = 2048                                         # CPU sample of the float32 DEM (same formula, top-left crop)
        r = synth.dem_float32(side * n_tiles, side, device=device).cpu().numpy()
        wins = [np.ascontiguousarray(r[:, i * side:(i + 1) * side]) for i in range(n_tiles)]
    return wins, level


def run_cpu(workload: str, n_tiles: int, threads: int, device, repeats: int = 1):
    """Returns (encode GS/s, decode GS/s, sample description)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import flac_oracle as fo
    fo.lib()
    wins, level = sample_windows(workload, n_tiles, device)
    jobs = [(w, level) for w in wins]
    best_e = best_d = 0.0
    for _ in range(repeats):
        t0 = time.perf_counter()
        if threads > 1:
            with ThreadPoolExecutor(threads) as ex:
                res = list(ex.map(cpu_tile_encode, jobs))
        else:
            res = [cpu_tile_encode(j) for j in jobs]
        te = time.perf_counter() - t0
        samples = sum(r[0] for r in res)
        from oracle import normalization_oracle as no
        djobs = []
        for (w, _), r in zip(jobs, res):
            bits = 16 if w.dtype.itemsize <= 2 else 24
            djobs.append((r[2], str(w.dtype), float(np.nanmin(w)), float(np.nanmax(w)), 32767 if bits == 16 else 8388607))
        t0 = time.perf_counter()
        if threads > 1:
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(cpu_tile_decode, djobs))
        else:
            [cpu_tile_decode(j) for j in djobs]
        td = time.perf_counter() - t0
        best_e = max(best_e, samples / te / 1e9)
        best_d = max(best_d, samples / td / 1e9)
    desc = f"first {n_tiles} tiles of the workload ({samples} samples), oracle C port of libFLAC 1.4.3 procedure + numpy normalisation"
    return best_e, best_d, desc, samples


_REAL_STDOUT = None


def emit_line(obj):
    """The ONE JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    # Libraries (NCCL's version banner, torch warnings) sometimes write to fd 1: route everything except the
    # final JSON line to stderr so stdout carries exactly one line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--scale-div", type=int, default=1, help="shrink the workload (debug only; invalid as a result)")
    ap.add_argument("--cpu-tiles", type=int, default=0, help="tiles in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    desc, level, tile_size, bands = WORKLOADS[args.workload]

    import torch

    if args.impl == "reference":
        if rank != 0:
            return 0
        dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
        cores = os.cpu_count() or 1
        n_tiles = args.cpu_tiles or max(2 * cores, 8)
        for _ in range(max(0, min(args.warmup, 1))):
            run_cpu(args.workload, min(n_tiles, cores), cores, dev)
        vals = []
        t_all = time.perf_counter()
        for _ in range(args.steps):
            e, d, sdesc, samples = run_cpu(args.workload, n_tiles, cores, dev)
            vals.append((e, d))
        ms = (time.perf_counter() - t_all) / max(args.steps, 1) * 1e3
        e = float(np.median([v[0] for v in vals]))
        d = float(np.median([v[1] for v in vals]))
        line = {
            "impl": "reference", "metric": METRIC, "value": e, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": desc, "level": level, "tile_size": tile_size, "blocksize": 4096},
            "decode": {"value": d, "unit": UNIT},
            "cpu_baseline": {"value": e, "unit": UNIT, "cores": cores, "kind": "port", "sample": sdesc,
                             "decode_value": d},
            "e2e": {"value": e, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit_line(line)
        return 0

    # ------------------------------------------------------------------ our arm
    from flac_raster_b200 import _native as nat
    from flac_raster_b200.distributed import allgather_tile_sizes, exclusive_scan, init_from_env
    from flac_raster_b200.engine import Engine, tile_grid
    import torch.distributed as dist

    nat.require_cuda()
    numa_node = None
    if world > 1:
        from flac_raster_b200.distributed import bind_to_gpu_numa_node
        numa_node = bind_to_gpu_numa_node(local)            # before any pinned allocation (first touch)
        print(f"[bench] rank {rank}: bound to NUMA node {numa_node}", file=sys.stderr)
        init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    eng = Engine(dev)
    L = nat.lib()
    L.frb_profile_enable(1)

    raster = make_raster(args.workload, dev, args.scale_div)
    nb, H, W = raster.shape
    ts = tile_size // (args.scale_div if args.workload != "c5" else 1)
    tiles = tile_grid(H, W, ts if args.workload != "c4" else max(H, W))
    n_tiles_local = len(tiles)
    samples_local = nb * H * W
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = {}

    def encode_step():
        enc = eng.encode_tiles(raster, tiles, level)
        if world > 1:        # the one collective of the path: global byte offsets of every tile
            sizes_all = _gather_equal(enc.sizes, world, dev)
            state["offsets_all"] = exclusive_scan(sizes_all)
        state["enc"] = enc
        return enc

    def _gather_equal(sizes, world, dev):
        send = torch.from_numpy(np.asarray(sizes, dtype=np.int64)).to(dev)
        recv = torch.empty(len(sizes) * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(recv, send)
        return recv.cpu().numpy()

    for _ in range(args.warmup):
        encode_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.frb_launch_count()
    enc_prof, dec_prof = {}, {}

    def prof(acc, slots):
        """Device time of the library's bracketed kernels for the step that just ran (CUDA events on its stream)."""
        for name, which in slots.items():
            ms = nat.C.c_float(0)
            if L.frb_profile_last_ms(which, nat.C.byref(ms)) == 0:
                acc.setdefault(name, []).append(ms.value)

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        encode_step()
        prof(enc_prof, {"k_enc_code": 0, "k_enc_stats": 4, "k_emit_frames": 2, "subframe_analysis_total": 6})
    ev1.record()
    barrier()
    launches = L.frb_launch_count() - launches0
    enc_ms = ev0.elapsed_time(ev1) / args.steps
    enc = state["enc"]
    comp_bytes = int(enc.sizes.sum())

    # ---- decode direction ------------------------------------------------------------------
    payload = torch.cat([enc.payload, torch.zeros(64, dtype=torch.uint8, device=dev)])
    out = torch.zeros(raster.numel() * raster.element_size(), dtype=torch.uint8, device=dev).view(raster.dtype).reshape(raster.shape)
    scale = 32767.0 if enc.bits_per_sample == 16 else 8388607.0

    def decode_step():
        # frames -> raster in one fused launch (sync scan, then skim + Rice decode + predictor restore + denormalise; CRC-16
        # beside it); the status read-back is part of the step
        return eng.decode_tiles(payload, enc.offsets, enc.sizes, tiles, enc.sample_rates, enc.minmax, scale, out, enc.bps, enc.blocksize)

    for _ in range(args.warmup):
        st = decode_step()
    assert list(st[:3]) == [0, 0, 0], f"decode status {st}"
    lossless = bool(torch.equal(out.reshape(-1).view(torch.uint8), raster.reshape(-1).view(torch.uint8))) if enc.bits_per_sample == 16 else None
    barrier()
    dlaunch0 = L.frb_launch_count()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for _ in range(args.steps):
        decode_step()
        prof(dec_prof, {"k_decode_subframes": 1, "k_skim_subframes": 5, "k_sync_scan": 3})
    ev3.record()
    barrier()
    dec_launches = L.frb_launch_count() - dlaunch0
    dec_ms = ev2.elapsed_time(ev3) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host (pinned) raster in, host frames out -----------------------------------------
    host_in = torch.empty(raster.numel() * raster.element_size(), dtype=torch.uint8).pin_memory()
    host_in.copy_(raster.reshape(-1).view(torch.uint8))
    host_raster = host_in.view(raster.dtype).reshape(raster.shape)
    host_out = torch.empty(comp_bytes + (1 << 20), dtype=torch.uint8).pin_memory()

    def e2e_step():
        # public engine call with HOST buffers: H2D of every tile row, encode, D2H of the frames (pipelined inside)
        e = eng.encode_tiles_host(host_raster, tiles, level, host_out=host_out)
        return int(e.payload.numel())

    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev4.record()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        nout = e2e_step()
    ev5.record()
    barrier()
    e2e_ms = max(ev4.elapsed_time(ev5), (time.perf_counter() - t0) * 1e3) / e2e_steps

    # ---- decode e2e: host (pinned) frames in, host raster out ---------------------------------------
    host_payload = torch.empty(comp_bytes, dtype=torch.uint8).pin_memory()
    host_payload.copy_(enc.payload[:comp_bytes])
    host_back = torch.empty(raster.numel() * raster.element_size(), dtype=torch.uint8).pin_memory().view(raster.dtype).reshape(raster.shape)

    def dec_e2e_step():
        return eng.decode_tiles_host(host_payload, enc.offsets, enc.sizes, tiles, enc.sample_rates, enc.minmax, scale, host_back,
                                     enc.bps, enc.blocksize)

    dec_e2e_ms = None
    if nb != 2:
        for _ in range(min(args.warmup, 2)):
            dec_e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            dst_ = dec_e2e_step()
        torch.cuda.synchronize()
        dec_e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        assert list(dst_[:3]) == [0, 0, 0], dst_
        barrier()

    # ---- reduce over ranks (max time) -------------------------------------------------------------
    t = torch.tensor([enc_ms, dec_ms, e2e_ms, dec_e2e_ms or 0.0], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(samples_local), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    enc_ms, dec_ms, e2e_ms, dec_e2e_ms = (float(v) for v in t.cpu())
    total_samples = float(tot[0])
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    traffic = {}
    tf = ROOT / "profiles" / "traffic.json"          # dram__bytes_read+write per launch from the committed ncu --set full captures
    if tf.exists():
        traffic = json.loads(tf.read_text()).get(args.workload, {})

    def roof(kernel_ms_list, alg_bytes, kernel, note):
        if not kernel_ms_list:
            return None
        kms = float(np.mean(kernel_ms_list))
        ach = alg_bytes / (kms * 1e-3) / 1e9
        extra = {k[len(kernel) + 1:]: v for k, v in traffic.items() if k.startswith(kernel + ":")} if args.scale_div == 1 else {}
        return {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": traffic.get(kernel) if args.scale_div == 1 else None, "kernel_ms": kms, **extra,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src, "note": note}

    def mean_ms(acc):
        return {k: float(np.mean(v)) for k, v in acc.items()}

    enc_alg = samples_local * 4 + comp_bytes           # int32 audio read + compressed bytes produced
    fused_dec = enc.bps == 16 and raster.element_size() <= 2 and nb != 2
    # compressed bytes consumed + output written: pixels of the raster (fused launch) or int32 audio (two-step path)
    dec_alg = comp_bytes + samples_local * (raster.element_size() if fused_dec else 4)
    cpu = None
    if not args.no_cpu_baseline and world == 1:      # the CPU baseline is reported by the single-GPU run only
        try:
            n_cpu = args.cpu_tiles or (8 if args.workload == "c3" else 64 if args.workload == "c5" else 4)
            ce, cd, sdesc, _ = run_cpu(args.workload, n_cpu, 1, dev)
            cpu = {"value": ce, "unit": UNIT, "cores": 1, "kind": "port", "sample": sdesc, "decode_value": cd}
        except Exception as ex:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"failed: {ex!r}"}

    line = {
        "metric": METRIC, "value": total_samples / (enc_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": enc_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": desc if args.scale_div == 1 else desc + f" [DEBUG scale-div {args.scale_div}]", "level": level,
                   "tile_size": ts, "blocksize": 4096, "tiles_per_gpu": n_tiles_local, "samples_per_gpu": samples_local,
                   "l2": "inputs (>= 1.9 GB per step) are far larger than the 126 MB L2; no flush needed",
                   "parallelism": f"tiles sharded over {world} GPU(s), one process per GPU; all-gather of per-tile sizes only",
                   "numa_bound_rank0": numa_node},
        "compressed_bytes_per_gpu": comp_bytes, "bits_per_sample_out": comp_bytes * 8 / samples_local,
        "lossless_roundtrip_checked": lossless,
        "decode": {"value": total_samples / (dec_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": dec_ms,
                   "gpu_launches": int(dec_launches),
                   "e2e": ({"value": total_samples / (dec_e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": dec_e2e_ms,
                            "h2d_bytes_per_step": comp_bytes, "d2h_bytes_per_step": int(raster.numel() * raster.element_size())}
                           if dec_e2e_ms else None),
                   "kernels_ms": mean_ms(dec_prof),
                   "roofline": roof(dec_prof.get("k_decode_subframes"), dec_alg, "k_decode_subframes",
                                    ("compressed bytes read + raster pixels written by the fused skim + Rice decode + predictor restore + "
                                     "denormalise launch" if fused_dec else "compressed bytes read + int32 audio written") +
                                    "; issue-bound (ALU pipe), not HBM-bound: see roofline.issue_* (ncu) and DESIGN.md section 4")},
        "kernels_ms": mean_ms(enc_prof),
        "roofline": roof(enc_prof.get("k_enc_code") or enc_prof.get("subframe_analysis_total"), enc_alg, "k_enc_code",
                         "int32 audio read + compressed bytes written by the dominant encode kernel (residual, Rice search, bit packing); "
                         "issue-bound, not HBM-bound (see DESIGN.md)"),
        "cpu_baseline": cpu,
        "e2e": {"value": total_samples / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(raster.numel() * raster.element_size()), "d2h_bytes_per_step": int(nout)},
        "gpu_launches": int(tot[1]),
        "clocks": clocks,
    }
    emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

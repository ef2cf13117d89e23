#!/usr/bin/env python3
"""bench.py -- decoded MP/s of the B200 JPEG decode path on the BASELINE.json workload.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): a batch of 4096
1920x1080 baseline 4:2:0 JPEGs per GPU, Pillow-encoded synthetic content (SURVEY.md 8d generator,
quality 75, no restart markers).  A "step" = one pass of the hot path over the whole batch:
device un-stuffing -> speculative Huffman decode -> fix-up/prefix -> coefficient write -> fused
IDCT/upsample/colour.  Images are independent: with N GPUs every rank decodes its own batch (weak
scaling, no collective on the data path; torch.distributed is only used for the timing barrier).

  value  whole-job MP/s with the compressed files already resident in HBM
  e2e    the same through the public pipeline object with the files in PINNED HOST memory: every step
         copies them host->device, runs all kernels, and reads the per-image status words back
  roofline / stages  per-kernel CUDA-event times, algorithmic bytes and GB/s against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (oracle/, a C port of the reference's algorithm) on one host core

`--impl reference` times the reference's algorithm on the host cores instead (oracle port, all threads).
"""
import argparse
import io
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W, H = 1920, 1080


def synth_image(seed: int, w: int = W, h: int = H) -> np.ndarray:
    """img = 128 + 100*(sin, cos, sin) + N(0, 12^2), seed = image index."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([128 + 100 * np.sin(x / 37 + y / 91), 128 + 100 * np.cos(x / 53 - y / 29),
                    128 + 100 * np.sin((x + y) / 71)], -1)
    img = img + rng.normal(0, 12, img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def encode_one(seed: int) -> bytes:
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(synth_image(seed)).save(b, "JPEG", quality=75, subsampling=2)
    return b.getvalue()


def make_files(n_distinct: int, seed0: int = 0):
    from multiprocessing import get_context
    workers = min(n_distinct, os.cpu_count() or 1, 16)
    if workers > 1:
        with get_context("fork").Pool(workers) as pool:
            return pool.map(encode_one, range(seed0, seed0 + n_distinct))
    return [encode_one(s) for s in range(seed0, seed0 + n_distinct)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args):
    """The reference's algorithm on the host cores: the CPU oracle (C port, oracle/) over all threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.lib()
    cores = os.cpu_count() or 1
    files = make_files(min(args.distinct, 2 * cores))
    per_step = max(cores, len(files))
    jobs = [files[i % len(files)] for i in range(per_step)]

    def one(d):
        return oracle.decode(d, want=("rgb",)).rgb.shape

    with ThreadPoolExecutor(cores) as ex:
        for _ in range(args.warmup):
            list(ex.map(one, jobs))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(ex.map(one, jobs))
        dt = time.perf_counter() - t0
    mp = per_step * args.steps * W * H / 1e6
    val = mp / dt
    line = {
        "impl": "reference", "metric": "decoded MP/s (1080p 4:2:0 batch)", "value": val, "unit": "MP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "1920x1080 baseline 4:2:0 q75 JPEGs, no restart markers (BASELINE.json configs[3])",
                   "images_per_step": per_step},
        "cpu_baseline": {"value": val, "unit": "MP/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} images/step x {args.steps} steps, oracle C port, {cores} threads"},
        "e2e": {"value": val, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=4096, help="images per GPU per step")
    ap.add_argument("--chunks", type=int, default=8, help="sub-batches per GPU, one CUDA stream each")
    ap.add_argument("--distinct", type=int, default=64, help="distinct synthetic images (cycled to --images)")
    ap.add_argument("--cpu-sample", type=int, default=48, help="images decoded by the 1-core CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # synthetic files first (fork pool before CUDA is touched); every rank gets its own seeds
    t_gen = time.perf_counter()
    files = make_files(args.distinct, seed0=rank * args.distinct)
    t_gen = time.perf_counter() - t_gen

    import torch
    import torch.distributed as dist
    from pyjpegdecoder_b200.parser import parse_jpeg
    from pyjpegdecoder_b200.pipeline import BatchPlan, DevicePipeline, pack_files, raise_for_errors

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_img = args.images
    n_chunks = max(1, min(args.chunks, n_img))
    parsed_d = [parse_jpeg(d) for d in files]
    # the batch is processed as n_chunks sub-batches, each with its own stream and buffers, so that the
    # host->device copy of one sub-batch and the latency-bound kernels of another overlap
    bounds = [round(i * n_img / n_chunks) for i in range(n_chunks + 1)]
    pipes, raws = [], []
    for c in range(n_chunks):
        idx = range(bounds[c], bounds[c + 1])
        datas = [files[i % len(files)] for i in idx]
        parsed = [parsed_d[i % len(files)] for i in idx]
        raw_host, offsets = pack_files(datas, pin=True)
        plan = BatchPlan(parsed, offsets, raw_host.numel())
        pipes.append(DevicePipeline(plan, dev, torch.cuda.Stream(dev)))
        raws.append(raw_host)
    main = torch.cuda.Stream(dev)
    scan_bytes = int(sum(int(p.plan.scans["raw_len"].sum()) for p in pipes))
    nblk = sum(p.plan.geom.total_blocks for p in pipes)
    n_sub = sum(int(p.plan.n_sub) for p in pipes)
    raw_bytes = sum(int(r.numel()) for r in raws)
    out_bytes = n_img * W * H * 3
    mp_per_step = n_img * W * H / 1e6

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(n_steps, body):
        """Run n_steps x body(pipe index) on the sub-batch streams, timed with CUDA events on `main`:
        every stream starts after e0 and e1 is recorded after all of them have finished."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record(main)
        for p in pipes:
            p.stream.wait_event(e0)
        for _ in range(n_steps):
            for c in range(len(pipes)):
                body(c)
        for p in pipes:
            ev = torch.cuda.Event()
            ev.record(p.stream)
            main.wait_event(ev)
        e1.record(main)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3 / n_steps
        return max(e0.elapsed_time(e1) / n_steps, 0.0), wall

    # ---- warm-up + correctness of the run ------------------------------------------------------------
    for p, r in zip(pipes, raws):
        p.upload(r)
    for _ in range(args.warmup):
        for p in pipes:
            p.launch()
    barrier()
    for p in pipes:
        raise_for_errors(p.err.cpu().numpy())

    # ---- per-stage times: sub-batches one after another, CUDA events around every stage --------------
    events = {}
    for _ in range(max(1, min(args.steps, 2))):
        for p in pipes:
            p.launch(events=events)
            p.stream.synchronize()
    stage_ms = {k: float(np.sum([a.elapsed_time(b) for (a, b) in v])) / max(1, min(args.steps, 2)) for k, v in events.items()}

    # ---- device-resident throughput (value): all sub-batch streams concurrently -----------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, _ = timed_steps(args.steps, lambda c: pipes[c].launch())
    ms_dev = max_over_ranks(ms_dev)

    # ---- end to end: pinned host bytes -> device -> kernels -> status words back ---------------------
    err_hosts = [torch.empty(len(p.plan.parsed), dtype=torch.int32, pin_memory=True) for p in pipes]

    def e2e_body(c):
        pipes[c].upload(raws[c])
        pipes[c].launch()
        with torch.cuda.stream(pipes[c].stream):
            err_hosts[c].copy_(pipes[c].err, non_blocking=True)

    timed_steps(max(1, args.warmup // 2), e2e_body)
    ms_ev, ms_wall = timed_steps(args.steps, e2e_body)
    ms_e2e = max_over_ranks(max(ms_ev, ms_wall))
    clocks = sampler.stop()
    assert all(int(e.abs().sum()) == 0 for e in err_hosts), "device reported decode errors"
    launches_per_step = sum(p.kernel_launches_per_step for p in pipes)
    device_bytes = sum(p.device_bytes() for p in pipes)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline accounting (algorithmic bytes) -------------------------------------
    peak, peak_src = measured_peak_gbs()
    alg = {
        "unstuff": 2 * scan_bytes,                       # read stuffed bytes, write compacted words
        "spec": scan_bytes + 32 * n_sub,                 # bitstream + entry/exit/count records
        "fix": 64 * n_sub,                               # records in, records + prefix out
        "write": scan_bytes + 128 * nblk,                # bitstream in, one 128-byte line per block out
        "pixels": 128 * nblk + out_bytes,                # coefficients in, RGB out (fused I+C)
    }
    stages = {}
    for k, ms in stage_ms.items():
        if k in alg and ms > 0:
            gbs = alg[k] / (ms * 1e-3) / 1e9
            stages[k] = {"ms": ms, "algorithmic_bytes": alg[k], "gbs": gbs, "frac_of_peak": gbs / peak,
                         "share_of_step": ms / sum(stage_ms.values())}
        else:
            stages[k] = {"ms": ms, "share_of_step": ms / sum(stage_ms.values())}
    dom = max((k for k in stages if "gbs" in stages[k]), key=lambda k: stages[k]["ms"])
    # dram__bytes_read.sum + dram__bytes_write.sum per image from the ncu --set full captures under profiles/
    # (r1g, 64 images per launch): measured DRAM traffic, scaled to the images one launch of this run covers
    ncu_traffic_per_image = {"pixels": (401.184512e6 + 354.579968e6) / 64, "write": (31.864320e6 + 348.049152e6) / 64,
                             "spec": 24.421376e6 / 64, "fix": (19.857408e6 + 3.593472e6) / 64,
                             "unstuff": (24.596224e6 + 24.657920e6) / 64}
    img_per_launch = n_img / n_chunks
    names = {"unstuff": "unstuff_count/scan_tiles/unstuff_scatter", "spec": "spec_kernel", "fix": "fix_local_kernel + chain_kernel",
             "write": "write_kernel", "pixels": "bj_pixels_fast_kernel<2,2,3> (fused dezigzag+dequant+IDCT+upsample+colour)"}

    def roof(k):
        st = stages.get(k, {})
        tr = ncu_traffic_per_image.get(k)
        return {"kernel": names[k], "bound": "hbm", "achieved": st.get("gbs"), "peak": peak, "unit": "GB/s",
                "frac": st.get("frac_of_peak"), "traffic": (tr * img_per_launch) if tr else None,
                "traffic_source": "ncu --set full capture profiles/r1g_full_summary.csv (64 images), scaled per image",
                "peak_source": peak_src, "launches_per_step": n_chunks,
                "algorithmic_bytes_per_launch": alg[k] / n_chunks, "ms_per_launch": st.get("ms", 0.0) / n_chunks,
                "share_of_step": st.get("share_of_step")}
    roofline = roof(dom)
    roofline_pixels = roof("pixels")

    # ---- the PUBLIC batch entry point, everything included (python bytes -> pinned pack, C marker walk +
    #      plan on the host, H2D, all kernels, status read-back); smaller batch, reported for information --
    api = None
    if world == 1:
        from pyjpegdecoder_b200.pipeline import decode_batch_on_device
        del pipes[:]
        torch.cuda.empty_cache()
        n_api = min(n_img, 1024)
        datas_api = [files[i % len(files)] for i in range(n_api)]
        for _ in range(2):
            decode_batch_on_device(datas_api, device=dev)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(3):
            decode_batch_on_device(datas_api, device=dev)
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / 3
        api = {"value": n_api * W * H / 1e6 / dt, "unit": "MP/s", "images": n_api, "ms": dt * 1e3,
               "what": "pyjpegdecoder_b200.pipeline.decode_batch_on_device(list of bytes): pack + host parse/plan + H2D + kernels + status"}

    # ---- CPU baseline: oracle port, one core, bounded sample ------------------------------------------
    cpu = None
    if world == 1:
        import oracle
        oracle.lib()
        sample = [files[i % len(files)] for i in range(args.cpu_sample)]
        t0 = time.perf_counter()
        for d in sample:
            oracle.decode(d, want=("rgb",))
        dt = time.perf_counter() - t0
        cpu = {"value": len(sample) * W * H / 1e6 / dt, "unit": "MP/s", "cores": 1, "kind": "port",
               "sample": f"{len(sample)} of the same 1080p files, oracle C port (oracle/jpeg_oracle.c), 1 thread, {dt:.1f} s"}

    line = {
        "metric": "decoded MP/s (1080p 4:2:0 batch)", "value": mp_per_step * world / (ms_dev * 1e-3), "unit": "MP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int16 (fp32 IDCT, fp64 tie-break)",
        "data": f"synthetic: {len(files)} distinct Pillow-encoded 1080p images per GPU cycled to {n_img}",
        "config": {"workload": "batch of 4096 1920x1080 baseline 4:2:0 q75 JPEGs per GPU, no restart markers "
                               "(BASELINE.json configs[3])",
                   "images_per_gpu": n_img, "sub_batches": n_chunks,
                   "streams": "one CUDA stream per sub-batch; stage times measured with the sub-batches run one after another",
                   "l2": "inputs larger than L2 (bitstream %.2f GB, coefficients %.1f GB per step)"
                         % (scan_bytes / 1e9, nblk * 128 / 1e9),
                   "host_parse": "marker parsing and descriptor build happen before the timed region"},
        "e2e": {"value": mp_per_step * world / (ms_e2e * 1e-3), "unit": "MP/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": raw_bytes, "d2h_bytes_per_step": 4 * n_img},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "roofline": roofline, "roofline_pixels": roofline_pixels, "stages": stages,
        "bitstream_gbs": {k: scan_bytes / (stage_ms[k] * 1e-3) / 1e9 for k in ("unstuff", "spec", "write") if k in stage_ms},
        "cpu_baseline": cpu, "public_api_e2e": api,
        "device_bytes": device_bytes, "gen_seconds": t_gen,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

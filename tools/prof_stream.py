"""Where the time goes in decode_stream: per-chunk GPU durations and gaps (CUDA events), host phases (wall clock)."""
import sys, time
sys.path.insert(0, '.')
from concurrent.futures import ThreadPoolExecutor
import bench, torch
from pyjpegdecoder_b200.loader import _Uploader, _check, _chunks
from pyjpegdecoder_b200.pipeline import decode_batch_on_device
files = bench.make_files(32)
datas = [files[i % 32] for i in range(8192)]
for rep in range(3):
    up = _Uploader("cuda:0")
    it = _chunks(datas, 512)
    T = dict(wait=0.0, enq=0.0, check=0.0)
    evs = []
    t_all = time.perf_counter()
    with ThreadPoolExecutor(1) as worker:
        queue = []
        k = 0
        def refill():
            global k
            while len(queue) < 2:
                nxt = next(it, None)
                if nxt is None:
                    return
                queue.append(worker.submit(up.prepare, nxt, k)); k += 1
        refill()
        prev = None
        while queue:
            t0 = time.perf_counter()
            chunk_files, packed, plan, raw_dev, (ev, evd), desc = queue.pop(0).result()
            torch.cuda.current_stream(up.dev).wait_event(evd)
            t1 = time.perf_counter(); T["wait"] += t1 - t0
            refill()
            raw_dev.record_stream(torch.cuda.current_stream(up.dev))
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record()
            batch = decode_batch_on_device(None, device="cuda:0", packed=packed, plan=plan, check=False, raw_dev=raw_dev, raw_ready=ev, desc=desc)
            b.record(); evs.append((a, b))
            t2 = time.perf_counter(); T["enq"] += t2 - t1
            if prev is not None:
                _check(prev)
                T["check"] += time.perf_counter() - t2
            prev = batch
        _check(prev)
    torch.cuda.synchronize()
    tot = time.perf_counter() - t_all
    dur = [x.elapsed_time(y) for x, y in evs]
    gaps = [evs[i][1].elapsed_time(evs[i + 1][0]) for i in range(len(evs) - 1)]
    print(f"total {1e3*tot:.1f} ms", {k_: round(1e3 * v, 1) for k_, v in T.items()},
          "gpu per chunk", [round(d, 1) for d in dur[:6]], "gaps", [round(g, 1) for g in gaps[:6]])

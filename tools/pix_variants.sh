#!/bin/bash
# Rebuild the library with different pixel-kernel tuning switches and time the 512-image step with each.
# usage (on a GPU box): bash tools/pix_variants.sh "-DBJ_PIX_CTAS=1 -DBJ_PIX_LOCKSTEP=1" "-DBJ_PIX_CTAS=2 -DBJ_PIX_LOCKSTEP=0" ...
for v in "$@"; do
  echo "=== $v"
  BJ_NVCC_EXTRA="$v" timeout 300 python -c "from pyjpegdecoder_b200.build import build_native; build_native(force=True)" || exit 1
  timeout 200 python bench.py --images 512 --chunks 1 --steps 5 --cpu-sample 2 2>&1 | tail -1 > gpurun_out/v.json
  timeout 20 python tools/bench_summary.py gpurun_out/v.json 2>&1 | sed -n 2,3p
done
# leave the default build behind
timeout 300 python -c "from pyjpegdecoder_b200.build import build_native; build_native(force=True)"

#!/bin/bash
# Build the library with extra -D switches (on the GPU box) and print the stage times of a short bench run.
# usage: tools/pix_exp.sh "<label>:<nvcc extra flags>" ...      e.g.  tools/pix_exp.sh "base:" "norecompute:-DBJ_EXP_NO_RECOMPUTE"
for spec in "$@"; do
  label="${spec%%:*}"; flags="${spec#*:}"
  BJ_NVCC_EXTRA="$flags" python -m pyjpegdecoder_b200.build --force > /dev/null || { echo "$label: build failed"; continue; }
  echo "== $label [$flags]"
  python bench.py --images 512 --chunks 1 --steps 5 --warmup 3 --cpu-sample 2 2>&1 | python tools/bench_summary.py | sed -n 1,3p
done
python -m pyjpegdecoder_b200.build --force > /dev/null

#!/bin/bash
# Full multi-stream bench with different warm-up lengths of spec_kernel (GPU box).  usage: tools/warm_exp.sh 4096 3072 2048
for w in "$@"; do
  BJ_NVCC_EXTRA="-DBJ_WARM_BITS=$w" python -m pyjpegdecoder_b200.build --force > /dev/null || { echo "$w: build failed"; continue; }
  python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('warm', $w, 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {k: round(v['ms'],2) for k,v in d['stages'].items()})
"
done
python -m pyjpegdecoder_b200.build --force > /dev/null

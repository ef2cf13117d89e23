#!/bin/bash
# Full multi-stream bench with other subsequence / warm-up lengths (GPU box).  usage: tools/subseq_exp.sh "8192 4096" "8192 8192"
for sw in "$@"; do
  set -- $sw
  BJ_NVCC_EXTRA="-DBJ_SUBSEQ_BITS=$1 -DBJ_WARM_BITS=$2" python -m pyjpegdecoder_b200.build --force > /dev/null || { echo "$sw: build failed"; continue; }
  timeout 300 python -m pytest tests/test_configs_gpu.py -x -q -m gpu -k "config1 or config2 or config4 or mixed_batch" 2>&1 | tail -1
  python bench.py --steps 5 --warmup 3 --cpu-sample 0 --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('subseq/warm', '$sw', 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {k: round(v['ms'],2) for k,v in d['stages'].items()})
"
done
python -m pyjpegdecoder_b200.build --force > /dev/null

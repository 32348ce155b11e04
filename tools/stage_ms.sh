#!/bin/bash
# per-stage tap-GEMM times of the 1080p inference plan under a few environment variants: tools/stage_ms.sh "VAR=1 VAR2=0" ...
for v in "$@"; do
  echo "== $v"
  env $v python bench.py --no-train --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fps %.1f one-lane %.1f' % (d['value'], d['one_lane']['value']))
print(' '.join('%s=%.3f' % (k.replace('.conv','c'),v) for k,v in d['stage_ms'].items() if not k.startswith('res') or k=='res1.conv1'))
"
done

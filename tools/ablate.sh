#!/bin/bash
# experiment helper: stage timings under VST_TG_DBG switches
for D in "$@"; do
  VST_TG_DBG=$D python bench.py --steps 6 --warmup 3 --frames-per-step 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('dbg', $D, round(d['value'],1), 'conv1', s['conv1'], 'conv2', s['conv2'], 'conv3', s['conv3'], 'trunk', s['res1.conv1'], 'deconv1', s['deconv1'], 'deconv2', s['deconv2'], 'deconv3', s['deconv3'])"
done

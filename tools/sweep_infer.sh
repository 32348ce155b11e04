#!/bin/bash
# experiment helper: inference-only bench lines under environment / flag variants, one compact line each
#   tools/sweep_infer.sh "ENV=.. ENV=.. -- --lanes 2 --frames-per-step 8" ...
for V in "$@"; do
  ENVS="${V%%--*}"; FLAGS="${V#*--}"
  env $ENVS python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --no-train $FLAGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms']
print('%-70s fps %.1f 1lane %.1f e2e %.1f | conv1 %.3f conv2 %.3f conv3 %.3f trunk %.3f dc1 %.3f dc2 %.3f dc3 %.3f | gemm share %.2f clk %s' % ('''$V''', d['value'], d['one_lane']['value'], d['e2e']['value'], s['conv1'], s['conv2'], s['conv3'], s['res1.conv1'], s['deconv1'], s['deconv2'], s['deconv3'], d['tapgemm_share_of_step'], d['clocks']['sm_mhz']))"
done

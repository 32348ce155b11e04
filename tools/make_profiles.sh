#!/bin/bash
# Round-2 profile artefacts (run on the GPU box; outputs under gpurun_out/, converted into profiles/ by tools/profiles_to_text.sh):
#  launch lists (gpu__time_duration + DRAM bytes, --clock-control none) of one inference forward and one training step,
#  ncu --set full captures of the trunk tap-GEMM, conv1, deconv2, the fused deconv3, an apply pass and warp_f32.
O=gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
python __graft_entry__.py --smoke > $O/r02_smoke.log 2>&1; tail -2 $O/r02_smoke.log
ncu --metrics $M --clock-control none -c 400 --csv --log-file $O/r02_launches_infer.csv python tools/one_forward.py 4 5 > $O/ncu_a.log 2>&1
ncu --metrics $M --clock-control none -c 1500 --csv --log-file $O/r02_launches_train.csv python tools/train_time.py bf16 short > $O/ncu_e.log 2>&1
cap() { # name, kernel regex, skip
  ncu --set full --import-source on --clock-control none -k regex:$2 -s $3 -c 1 -o $O/r02_$1 -f python tools/one_forward.py 4 2 > $O/ncu_$1.log 2>&1
}
cap trunk_tapgemm tapgemm_kernel 19
cap conv1_tapgemm tapgemm_kernel 16
cap deconv2_tapgemm tapgemm_kernel 30
cap deconv3_fused_tapgemm tapgemm_kernel 31
cap apply_conv1 apply_lds 15
ncu --set full --clock-control none -k regex:warp_f32 -s 3 -c 1 -o $O/r02_warp_f32 -f python -c "
import sys; sys.path.insert(0, '.')
import torch, vst_b200
from vst_b200 import ops
g = torch.Generator('cuda').manual_seed(1)
x = torch.rand((8, 3, 1024, 1024), device='cuda', generator=g) * 255
f = torch.randn((8, 2, 1024, 1024), device='cuda', generator=g) * 4
for _ in range(6): ops.warp(x, f)
torch.cuda.synchronize()" > $O/ncu_warp.log 2>&1
ls -la $O/r02_*.ncu-rep $O/r02_launches_*.csv

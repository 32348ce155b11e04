#!/bin/bash
# gpurun_out/r02_* -> profiles/ (text summaries the judge can read without ncu)
O=${O:-gpurun_out}; P=${P:-profiles}
if [ $# -lt 4 ]; then set -- $(python tools/profile_ranges.py $O/r02_launches_infer.csv $O/r02_launches_train.csv); fi
hdr() { echo "# $1"; echo "# command: $2"; }
{ hdr "ncu launch list, one ReCoNet 1080p forward of 4 frames on one lane (round 2, final: staged epilogue in ping-pong on the 48-channel layers, deconv3 with fused input normalisation + tap-merged MMAs, programmatic dependent launch); times are cold-cache and serialised" "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv python tools/one_forward.py 4 5";
  python tools/ncu_summary.py $O/r02_launches_infer.csv $1 $2; echo; python tools/ncu_perlaunch.py $O/r02_launches_infer.csv $1 $2; } > $P/r02_launches_infer_1080p_b4.txt
{ hdr "ncu launch list, one bf16 ReCoNet training step (1024x436, 2 pairs), eager launches (bench replays the same step as a CUDA graph)" "ncu --metrics ... -c 1500 --csv python tools/train_time.py bf16 short";
  python tools/ncu_summary.py $O/r02_launches_train.csv $3 $4; } > $P/r02_launches_train_1024x436_b2.txt
for n in trunk_tapgemm conv1_tapgemm deconv2_tapgemm deconv3_fused_tapgemm apply_conv1 warp_f32 gram_c64; do
  ncu -i $O/r02_$n.ncu-rep --page details 2>/dev/null | grep -v "^\s*$" > $P/r02_${n}_ncu_details.txt
done
python tools/ncu_traffic.py $P/r02_traffic.json trunk_tapgemm=$O/r02_trunk_tapgemm.ncu-rep:4 conv1_tapgemm=$O/r02_conv1_tapgemm.ncu-rep:4 deconv2_tapgemm=$O/r02_deconv2_tapgemm.ncu-rep:4 deconv3_fused_tapgemm=$O/r02_deconv3_fused_tapgemm.ncu-rep:4 apply_conv1=$O/r02_apply_conv1.ncu-rep:4 > /dev/null
cuobjdump -sass video-style-transfer_b200/csrc/libvst_b200.so > /tmp/sass.txt
{ echo "# SASS evidence, csrc/libvst_b200.so (round 2)"; for k in UTCHMMA UTCHMMA.2CTA UTMALDG LDTM UTCBAR "STG.E.ENL2.256" SHFL.BFLY "REDG.E.ADD.F64"; do echo "$k: $(grep -c "$k" /tmp/sass.txt)"; done; } > $P/r02_sass_evidence.txt
cat $P/r02_sass_evidence.txt

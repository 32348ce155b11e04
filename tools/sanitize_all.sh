#!/bin/bash
# compute-sanitizer over every tap-GEMM / pixel-contraction-GEMM mode; usage: tools/sanitize_all.sh memcheck|racecheck|synccheck OUTDIR
TOOL=$1; OUT=$2; mkdir -p $OUT
run() { # name, what, env...
  local name=$1 what=$2; shift 2
  env "$@" timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_run.py $what > $OUT/${TOOL}_$name.log 2>&1
  echo "$TOOL $name rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok' $OUT/${TOOL}_$name.log | tr '\n' ' ')"
}
run default          all   VST_X=0
run per_tap_boxes    infer VST_STREAM=0 VST_DYSHARE=0 VST_CTA2=0
run dyshare_only     infer VST_STREAM=0 VST_DYSHARE=1 VST_CTA2=0
run cta_pair_only    infer VST_STREAM=0 VST_DYSHARE=0 VST_CTA2=1
run row_ring         infer VST_STREAM=1 VST_DYSHARE=0
run acc_ring         infer VST_STREAM=2 VST_DYSHARE=0
run staged_epilogue  all   VST_EPI_DIRECT=0
run pcgemm_plain     train VST_PC_MCHUNK=0

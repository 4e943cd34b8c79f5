#!/bin/bash
# In-kernel timelines of the fused blocks: variants lib_tl104.so (rb4) / lib_tl101.so (rb1) built with
# TDM_NVCC_DEFS=-DTDM_TIMELINE=<104|101> and copied to tinydiffusionmodels_b200/build/variants/.
cd "$(dirname "$0")/.."
cp tinydiffusionmodels_b200/libtdm_b200.so /tmp/lib_keep.so
for k in 104 101; do
  cp tinydiffusionmodels_b200/build/variants/lib_tl$k.so tinydiffusionmodels_b200/libtdm_b200.so
  timeout 200 python tools/fused_timeline.py 16384 > gpurun_out/tl_$k.log 2>&1
done
cp /tmp/lib_keep.so tinydiffusionmodels_b200/libtdm_b200.so

#!/bin/bash
# Time one fused p_sample per library variant under tinydiffusionmodels_b200/build/variants/
# (variants are built here with TDM_NVCC_DEFS=... python tinydiffusionmodels_b200/build.py and copied there).
cd "$(dirname "$0")/.."
cp tinydiffusionmodels_b200/libtdm_b200.so /tmp/lib_keep.so
for v in tinydiffusionmodels_b200/build/variants/lib_*.so; do
  cp "$v" tinydiffusionmodels_b200/libtdm_b200.so
  echo "== $(basename $v)"
  timeout 200 python tools/perf_probe.py ${SWEEP_BATCHES:-16384} 2>&1 | tail -1
done
cp /tmp/lib_keep.so tinydiffusionmodels_b200/libtdm_b200.so

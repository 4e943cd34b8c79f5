#!/bin/bash
# Text reverse-step time per library variant under tinydiffusionmodels_b200/build/variants/.
cd "$(dirname "$0")/.."
cp tinydiffusionmodels_b200/libtdm_b200.so /tmp/lib_keep.so
echo "== default"; timeout 200 python tools/perf_text.py 2>&1 | grep "text p_sample dim=256 B=\(512\|2048\)"
for v in tinydiffusionmodels_b200/build/variants/lib_*.so; do
  cp "$v" tinydiffusionmodels_b200/libtdm_b200.so
  echo "== $(basename $v)"
  timeout 200 python tools/perf_text.py 2>&1 | grep "text p_sample dim=256 B=\(512\|2048\)\|rror"
done
cp /tmp/lib_keep.so tinydiffusionmodels_b200/libtdm_b200.so

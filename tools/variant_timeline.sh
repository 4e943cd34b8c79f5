#!/bin/bash
# cycles per tile (in-kernel timeline, clock-independent) of each library variant under build/variants/
# (variants built with -DTDM_TIMELINE=104 or 101 plus whatever is being compared)
cd "$(dirname "$0")/.."
cp tinydiffusionmodels_b200/libtdm_b200.so /tmp/lib_keep.so
for v in tinydiffusionmodels_b200/build/variants/lib_*.so; do
  cp "$v" tinydiffusionmodels_b200/libtdm_b200.so
  echo "== $(basename $v)"
  timeout 200 python tools/fused_timeline.py 16384 2>&1 | tail -4
done
cp /tmp/lib_keep.so tinydiffusionmodels_b200/libtdm_b200.so

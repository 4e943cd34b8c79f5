#!/bin/bash
# A/B of library variants under tinydiffusionmodels_b200/build/variants/ in both regimes: per-kernel event times of a
# cold step (perf_probe.py) and us per reverse step of full T=1000 trajectories (hot_probe.py, the power-capped regime).
cd "$(dirname "$0")/.."
cp tinydiffusionmodels_b200/libtdm_b200.so /tmp/lib_keep.so
for v in tinydiffusionmodels_b200/build/variants/lib_*.so; do
  cp "$v" tinydiffusionmodels_b200/libtdm_b200.so
  echo "== $(basename $v)"
  timeout 200 python tools/perf_probe.py ${SWEEP_BATCHES:-16384} 2>&1 | tail -1
  timeout 300 python tools/hot_probe.py ${HOT_BATCH:-16384} 2 2>&1 | tail -1
done
cp /tmp/lib_keep.so tinydiffusionmodels_b200/libtdm_b200.so

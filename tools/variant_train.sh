#!/bin/bash
# Training throughput (bench.py's train leg) per library variant under tinydiffusionmodels_b200/build/variants/.
cd "$(dirname "$0")/.."
cp tinydiffusionmodels_b200/libtdm_b200.so /tmp/lib_keep.so
for rep in 1 2; do
for v in tinydiffusionmodels_b200/build/variants/lib_*.so; do
  cp "$v" tinydiffusionmodels_b200/libtdm_b200.so
  echo "== $(basename $v)"
  timeout 300 python bench.py --no-text --batch 2048 --steps 1 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train img/s', round(d['train']['value']), 'ms', d['train'].get('ms_per_step'))"
done
done
cp /tmp/lib_keep.so tinydiffusionmodels_b200/libtdm_b200.so

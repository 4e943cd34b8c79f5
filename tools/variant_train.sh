#!/bin/bash
# Training throughput (bench.py's train leg) per library variant under tinydiffusionmodels_b200/build/variants/,
# alternating the variants twice so box drift shows up.   TRAIN_BATCHES="512 8192" bash tools/variant_train.sh
cd "$(dirname "$0")/.."
cp tinydiffusionmodels_b200/libtdm_b200.so /tmp/lib_keep.so
for rep in 1 2; do
for v in tinydiffusionmodels_b200/build/variants/lib_*.so; do
  cp "$v" tinydiffusionmodels_b200/libtdm_b200.so
  for b in ${TRAIN_BATCHES:-512}; do
    timeout 300 python bench.py --no-text --batch 1024 --steps 1 --warmup 3 --train-batch $b 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$(basename $v) batch $b: img/s', round(d['train']['value']), 'ms', round(d['train'].get('ms_per_step'),4))"
  done
done
done
cp /tmp/lib_keep.so tinydiffusionmodels_b200/libtdm_b200.so

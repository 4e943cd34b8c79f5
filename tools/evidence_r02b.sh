#!/bin/bash
# Round-2 evidence for the text GEMMs and the Shakespeare training step (row f2).  One gpurun call; every ncu run follows a
# plain run of the same command that exited 0.   bash tools/evidence_r02b.sh  -> gpurun_out/r02b_*
cd "$(dirname "$0")/.."
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
# (1) launch list of the training step at the reference CLI's batch
python tools/text_train_probe.py 32 > gpurun_out/r02b_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02b_launches_text_train_raw.csv \
    python tools/text_train_probe.py 32 > gpurun_out/r02b_ncu1.log 2>&1
# (2) per-kernel metrics of the step's large kernels (one step: skip the first step's launches)
ncu --metrics $M --clock-control none -k regex:'gemm_tc_kernel<\(int\)[56]>|adamw_vec|plane_colsum' -s 6 -c 6 --csv \
    --log-file gpurun_out/r02b_text_train_metrics_raw.csv python tools/text_train_probe.py 32 > gpurun_out/r02b_ncu2.log 2>&1
# (3) the rounding GEMM + argmax and the guided mix
python tools/perf_text.py > gpurun_out/r02b_plain3.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:'gemm_tc_kernel<\(int\)2>' --csv --log-file gpurun_out/r02b_rounding_metrics_raw.csv \
    python tools/perf_text.py > gpurun_out/r02b_ncu3.log 2>&1
ls -la gpurun_out/r02b_*raw.csv

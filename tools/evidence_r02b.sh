#!/bin/bash
# Round-2 evidence for the text GEMMs and the Shakespeare training step (row f2).  One gpurun call; every ncu run follows a
# plain run of the same command that exited 0.   bash tools/evidence_r02b.sh  -> gpurun_out/r02b_*
cd "$(dirname "$0")/.."
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
# (1) launch list of the training step at the reference CLI's batch   (SKIP1=1 skips it)
python tools/text_train_probe.py 32 > gpurun_out/r02b_plain1.log 2>&1 && [ -z "$SKIP1" ] &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02b_launches_text_train_raw.csv \
    python tools/text_train_probe.py 32 > gpurun_out/r02b_ncu1.log 2>&1
# (2) per-kernel metrics of the step's GEMMs (the second step's 40 launches: 12 forward, log-sum-exp, d logits, dX0, dW, 24 backward)
#     and of its HBM-bound kernels
ncu --metrics $M --clock-control none -k regex:'gemm_tc_kernel' -s 40 -c 40 --csv \
    --log-file gpurun_out/r02b_text_train_gemm_metrics_raw.csv python tools/text_train_probe.py 32 > gpurun_out/r02b_ncu2.log 2>&1
ncu --metrics $M --clock-control none -k regex:'adamw_vec|plane_colsum' -s 2 -c 2 --csv \
    --log-file gpurun_out/r02b_text_train_hbm_metrics_raw.csv python tools/text_train_probe.py 32 > gpurun_out/r02b_ncu2b.log 2>&1
# (3) the rounding GEMM + argmax (32,768 and 4,096 rows) and the guided mix (512 sequences)
python tools/round_step.py > gpurun_out/r02b_plain3.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:'gemm_tc_kernel' --csv --log-file gpurun_out/r02b_rounding_metrics_raw.csv \
    python tools/round_step.py > gpurun_out/r02b_ncu3.log 2>&1
ls -la gpurun_out/r02b_*raw.csv

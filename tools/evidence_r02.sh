#!/bin/bash
# Round-2 ncu evidence (one gpurun call; every ncu run follows a plain run of the same command that exited 0).
#   bash tools/evidence_r02.sh            -> gpurun_out/r02_*.csv ; summarise here with tools/traffic_step.py etc.
cd "$(dirname "$0")/.."
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
# (1) per-kernel metrics of the fused reverse step at 16,384 images
python tools/one_step.py 16384 2 > gpurun_out/ev_plain1.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:'conv3x3|resblock|avgpool' --csv --log-file gpurun_out/r02_step_metrics_raw.csv \
    python tools/one_step.py 16384 2 > gpurun_out/ev_ncu1.log 2>&1
# (2) the elementwise kernels
python tools/elementwise_step.py > gpurun_out/ev_plain2.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
    --clock-control none -k regex:'q_sample|reverse_step|randn|unit_range' --csv --log-file gpurun_out/r02_elementwise_raw.csv \
    python tools/elementwise_step.py > gpurun_out/ev_ncu2.log 2>&1
# (3) launch list of the bench command (shares of the step)
python bench.py --steps 1 --warmup 3 --no-text --no-extras > gpurun_out/ev_plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r02_launches_bench_raw.csv \
    python bench.py --steps 1 --warmup 3 --no-text --no-extras > gpurun_out/ev_ncu3.log 2>&1
# (4) one text reverse step and one training step
python tools/text_step.py 512 > gpurun_out/ev_plain4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_text_raw.csv \
    python tools/text_step.py 512 > gpurun_out/ev_ncu4.log 2>&1
python tools/train_step.py 512 > gpurun_out/ev_plain5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train_raw.csv \
    python tools/train_step.py 512 > gpurun_out/ev_ncu5.log 2>&1
ls -la gpurun_out/r02_*raw.csv

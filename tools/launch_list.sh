#!/bin/bash
# ncu launch list (gpu__time_duration.sum per launch) of a command, after the same command exited 0 without ncu.
#   bash tools/launch_list.sh <out.csv> <command...>        (run through gpurun; one ncu per gpurun call)
out="$1"; shift
timeout 300 "$@" > /dev/null 2>&1 || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out" "$@" > /dev/null 2>&1
wc -l "$out"

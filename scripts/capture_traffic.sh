#!/bin/bash
# Round-2 ncu captures behind profiles/r02_traffic.json (run on the GPU box: gpurun -- 'bash scripts/capture_traffic.sh trace|render').
# Each ncu run is preceded by the same command without ncu (B200_PROFILING.md).
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
if [ "$1" = "trace" ]; then
  python scripts/trace_speed.py --reps 1 > gpurun_out/ts_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:trace_g2 -s 2 -c 1 -o gpurun_out/r02_prof_trace_headline python scripts/trace_speed.py --reps 1 > gpurun_out/ts_ncu.log 2>&1
  python scripts/trace_speed.py --reps 1 --lbvh > gpurun_out/tsl_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:trace_g2 -s 2 -c 1 -o gpurun_out/r02_prof_trace_lbvh python scripts/trace_speed.py --reps 1 --lbvh > gpurun_out/tsl_ncu.log 2>&1
  tail -n 1 gpurun_out/ts_plain.log gpurun_out/tsl_plain.log
else
  for c in "5 4 " "5 4 --lbvh" "3 16 " "4 16 " "1 64 "; do
    set -- $c
    tag=cfg$1$( [ -n "$3" ] && echo _lbvh )
    K="-k regex:extend_"
    [ "$1" != "5" ] && K=""
    python scripts/render_one.py --config $1 --spp $2 $3 --no-warmup --stats > gpurun_out/r02_cap_${tag}_plain.log 2>&1 && \
    ncu --metrics $M --clock-control none $K --csv --log-file gpurun_out/r02_cap_${tag}.csv python scripts/render_one.py --config $1 --spp $2 $3 --no-warmup --stats > gpurun_out/r02_cap_${tag}_ncu.log 2>&1
    tail -n 1 gpurun_out/r02_cap_${tag}_plain.log
  done
fi
echo done

M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
for c in "4 16" "1 64"; do
  set -- $c
  python scripts/render_one.py --config $1 --spp $2 --no-warmup --stats > gpurun_out/r02_cap_cfg$1_plain.log 2>&1 && \
  timeout 200 ncu --metrics $M --clock-control none -k regex:extend_ --csv --log-file gpurun_out/r02_cap_cfg$1.csv python scripts/render_one.py --config $1 --spp $2 --no-warmup --stats > gpurun_out/r02_cap_cfg$1_ncu.log 2>&1
  tail -n 1 gpurun_out/r02_cap_cfg$1_plain.log | cut -c1-400
done

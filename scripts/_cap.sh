M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
for c in "3 16" "4 16" "1 64"; do
  set -- $c
  tag=cfg$1
  python scripts/render_one.py --config $1 --spp $2 --no-warmup --stats > gpurun_out/r02_cap_${tag}_plain.log 2>&1 && \
  ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_cap_${tag}.csv python scripts/render_one.py --config $1 --spp $2 --no-warmup --stats > gpurun_out/r02_cap_${tag}_ncu.log 2>&1
  tail -n 1 gpurun_out/r02_cap_${tag}_plain.log
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/r02_bench_norender.json 2> gpurun_out/plain_bench.log && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/ncu_bench.json 2> gpurun_out/ncu_bench.log
echo done

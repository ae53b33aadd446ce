python scripts/render_one.py --config 5 --spp 4 --no-warmup > gpurun_out/plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:extend_g2 -c 2 -o gpurun_out/r02_prof_cfg5_extend python scripts/render_one.py --config 5 --spp 4 --no-warmup > gpurun_out/ncu5.log 2>&1
python scripts/render_one.py --config 4 --spp 4 --no-warmup > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"shade_kernel|extend_kernel" -c 8 -o gpurun_out/r02_prof_cfg4_after python scripts/render_one.py --config 4 --spp 4 --no-warmup > gpurun_out/ncu4.log 2>&1
echo done

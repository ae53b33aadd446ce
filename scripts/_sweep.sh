python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "same_path or aov or tile or depth" 2>&1 | tail -3
for L in "" variants/lib_s6_d5.so variants/lib_s7_d5.so variants/lib_s8_d6.so; do
  echo "== $L"
  IZPI_LIB_PATH=$L python scripts/render_one.py --config 4 --spp 16 2>&1 | tail -1
  IZPI_LIB_PATH=$L python scripts/render_one.py --config 1 --spp 64 --repeat 3 2>&1 | tail -1
done

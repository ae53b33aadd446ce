python -m pytest tests/test_render_gpu.py tests/test_group_gpu.py -m gpu -x -q -k "same_path or aov or tile or depth or sorted or cursor or image" 2>&1 | tail -3
python scripts/render_one.py --config 4 --spp 64 2>&1 | tail -1
IZPI_LIB_PATH=variants/lib_plinline.so python scripts/render_one.py --config 4 --spp 64 2>&1 | tail -1
python scripts/render_one.py --config 4 --spp 64 2>&1 | tail -1
IZPI_LIB_PATH=variants/lib_plinline.so python scripts/render_one.py --config 4 --spp 64 2>&1 | tail -1

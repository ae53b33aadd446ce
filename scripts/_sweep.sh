python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "same_path or aov or tile or depth or sorted" 2>&1 | tail -3
python scripts/material_sort_ab.py --materials 8 --spp 32 2>&1 | tail -1
python scripts/render_one.py --config 3 --spp 32 2>&1 | tail -1
python scripts/render_one.py --config 4 --spp 64 2>&1 | tail -1
python scripts/render_one.py --config 1 --spp 64 --repeat 3 2>&1 | tail -1

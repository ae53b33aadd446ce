python -m pytest tests/test_render_gpu.py tests/test_group_gpu.py tests/test_proto_scene.py -m gpu -x -q -k "not converged" 2>&1 | tail -3
IZPI_DEFER_PROBE=0 python scripts/render_one.py --config 4 --spp 64 --repeat 3 2>&1 | tail -1
python scripts/render_one.py --config 4 --spp 64 --repeat 3 2>&1 | tail -1
python scripts/render_one.py --config 4 --spp 64 --repeat 2 --stats 2>&1 | tail -1
IZPI_SORT_RAYS=0 python scripts/render_one.py --config 4 --spp 64 --repeat 3 2>&1 | tail -1

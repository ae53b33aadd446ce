python -m pytest tests/test_trace_gpu.py tests/test_bvh_device.py -m gpu -x -q 2>&1 | tail -3
IZPI_RAYF_PREPASS=0 python scripts/trace_speed.py --reps 5 2>&1 | tail -1
python scripts/trace_speed.py --reps 5 2>&1 | tail -1
IZPI_RAYF_PREPASS=0 python scripts/trace_speed.py --reps 5 --lbvh 2>&1 | tail -1
python scripts/trace_speed.py --reps 5 --lbvh 2>&1 | tail -1

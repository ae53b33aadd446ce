for L in "" variants/lib_key0.so variants/lib_key2.so; do
echo "== $L"
IZPI_LIB_PATH=$L python scripts/render_one.py --config 4 --spp 64 --repeat 2 --stats 2>&1 | tail -1
IZPI_LIB_PATH=$L python scripts/render_one.py --config 4 --spp 64 --repeat 3 2>&1 | tail -1
done

python -m pytest tests -m gpu -x -q -k "lzss or golden or prefix or multi or batch" 2>&1 | tail -2
python tools/sweep_env.py AGMVB_LZ_REFILL 1 4 8 16 -- --steps 3 --warmup 2 | cut -c1-250

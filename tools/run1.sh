for o in 4 3; do python tools/quick.py --m 6 --n 20 --batch 1 --iters 20 --flush --check 0 --opt slab_ctas=$o 2>&1 | grep -E "ms/step|1601"; done
for o in 4 3; do python tools/quick.py --m 4 --n 20 --batch 64 --iters 20 --check 0 --opt slab_ctas=$o 2>&1 | grep -E "ms/step|1601"; done
python -m pytest tests -m gpu -x -q -k "golden or c3 or alternative" 2>&1 | tail -2

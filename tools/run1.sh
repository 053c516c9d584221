set -x
python -m pytest tests -m gpu -x -q -k "golden or alternative or property or generic or c3" 2>&1 | tail -3
for o in 1 2 0; do python tools/quick.py --m 6 --n 20 --batch 1 --iters 20 --flush --check 0 --opt use_cta_cap=$o; done
python tools/quick.py --m 4 --n 20 --batch 64 --iters 20 --check 0 --opt use_cta_cap=1 --trace 0
python tools/quick.py --m 4 --n 20 --batch 64 --iters 20 --check 0 --opt use_cta_cap=2 --trace 0

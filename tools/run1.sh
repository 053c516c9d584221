python tools/quick.py --m 6 --n 20 --batch 1 --iters 20 --flush --check 0 2>&1 | grep -E "ms/step|label"
python tools/quick.py --m 4 --n 20 --batch 64 --iters 20 --check 0 2>&1 | grep -E "ms/step|label"
python tools/quick.py --m 5 --n 12 --batch 16 --iters 20 --check 0 2>&1 | grep -E "ms/step"
python -m pytest tests -m gpu -x -q -k "golden or c3 or alternative or property" 2>&1 | tail -2

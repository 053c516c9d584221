python -m pytest tests -m gpu -x -q -k "whole or golden or push or odd or shapes" 2>&1 | tail -2
python tools/quick.py --m 3 --n 30 --batch 4096 --iters 10 2>&1 | tail -3
python tools/quick.py --m 3 --n 20 --batch 8192 --iters 10 2>&1 | tail -2
python tools/quick.py --m 2 --n 30 --batch 16384 --iters 10 2>&1 | tail -2

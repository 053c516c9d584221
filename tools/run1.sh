set -x
python -m pytest tests -m gpu -x -q -k "alternative or odd or golden or whole" 2>&1 | tail -3
python tools/quick.py --batch 4096 --iters 10 --trace 0
python tools/quick.py --batch 4096 --iters 10 --trace 0 --check 0 --sweep whole_order_exact=104,64,100,96,68,80
python tools/quick.py --batch 4096 --iters 10 --trace 0 --check 0 --sweep whole_order_fma=104,64,100,96,68,80

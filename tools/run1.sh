set -x
python tools/quick.py --batch 4096 --iters 10 --trace 0 --check 0 --colex 0 --opt whole_split=1 --sweep whole_split_skew=0,5000,10000,20000,30000,45000
python tools/quick.py --batch 4096 --iters 10 --trace 0 --check 0 --colex 1 --opt whole_split=1 --sweep whole_split_skew=0,10000,20000,30000

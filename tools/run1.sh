export LPFNT_DEBUG_OCCUPANCY=1
python tools/quick.py --batch 4096 --iters 10 --trace 0 --check 0 --colex 0 --opt whole_split=1 2>&1 | tail -4
python tools/quick.py --batch 4096 --iters 10 --trace 0 --check 0 --colex 1 --opt whole_split=1 2>&1 | tail -4
python tools/quick.py --batch 4096 --iters 10 --trace 0 --check 0 --colex 0 --opt whole_split=1 --sweep whole_split_c=11,12,13 2>&1 | tail -8

python tools/quick.py --m 3 --n 30 --batch 4096 --iters 10 --check 0 2>&1 | tail -1
python tools/quick.py --m 3 --n 30 --batch 4096 --iters 10 --check 0 --opt whole_threads_exact=384 --opt whole_fibres_exact=2 2>&1 | tail -1
python tools/quick.py --m 3 --n 30 --batch 4096 --iters 10 --check 0 --opt whole_threads_fma=768 --opt whole_fibres_fma=1 2>&1 | tail -1

set -x
python -m pytest tests -m gpu -x -q -k "alternative or whole or golden or odd" 2>&1 | tail -3
python tools/quick.py --batch 4096 --iters 10 --trace 0
python tools/quick.py --batch 4096 --iters 10 --trace 0 --check 0 --sweep whole_order_exact=40,104,0,64,42,36,20,21,100
python tools/quick.py --batch 4096 --iters 10 --trace 0 --check 0 --opt whole_fibres_exact=2 --sweep whole_order_exact=40,104,0,64,36
python tools/quick.py --batch 4096 --iters 10 --trace 0 --check 0 --opt whole_threads_exact=768 --opt whole_fibres_exact=1 --sweep whole_order_exact=40,104
python tools/quick.py --batch 4096 --iters 10 --trace 0 --check 0 --opt whole_fibres_fma=3 --sweep whole_order_fma=40,104,0,36
python tools/timeline.py 30 > gpurun_out/timeline_r2h.txt 2>&1; cat gpurun_out/timeline_r2h.txt | grep -E "==|period|warp  0|warp  1:|warp 1[12]|warp 23| warp  6"

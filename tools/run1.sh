set -x
python tools/quick.py --batch 4096 --iters 10 --trace 0 --opt whole_threads_exact=640 --opt whole_fibres_exact=1
python tools/quick.py --batch 4096 --iters 10 --trace 0 --check 0 --opt whole_threads_exact=640 --opt whole_fibres_exact=1 --sweep whole_order_exact=40,104,36,100,0,64
python tools/quick.py --batch 4096 --iters 10 --trace 0 --check 0 --sweep whole_order_exact=104
python tools/timeline.py 30 whole_threads_exact=640 whole_fibres_exact=1 > gpurun_out/timeline_r2i.txt 2>&1; grep -E "==|period|warp  0|warp 19|warp 10" gpurun_out/timeline_r2i.txt | head -12

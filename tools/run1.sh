python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29549 tests/run_push_gpu.py 2>&1 | grep -E "^push|PUSH|rror|multicast|Multicast" | tail -20

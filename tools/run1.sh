cat > /tmp/d6.py <<'PY'
import sys; sys.path.insert(0, "/root/repo")
import torch, lpfun_b200
t = lpfun_b200.Transform(6, 20, report=False)
v = torch.randn(len(t), dtype=torch.float64, device="cuda"); c = torch.empty_like(v); r = torch.empty_like(v)
for _ in range(3):
    t.fnt(v, out=c); t.ifnt(c, out=r)
torch.cuda.synchronize()
PY
python /tmp/d6.py && ncu --set full --clock-control none --import-source on -k regex:"k_slab|k_sweep_groups" -s 10 -c 5 -o gpurun_out/prof_d6_r2 -f python /tmp/d6.py > gpurun_out/ncu_d6.log 2>&1
tail -2 gpurun_out/ncu_d6.log

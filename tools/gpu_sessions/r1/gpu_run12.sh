cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q --timeout 300 -k "eig or head_mds or vgru" > gpurun_out/r12_eig.log 2>&1
echo "eig exit $?" >> gpurun_out/r12_eig.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r12_all.log 2>&1
echo "all exit $?" >> gpurun_out/r12_all.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r12_bench.json 2> gpurun_out/r12_bench.err
timeout 300 python - > gpurun_out/r12_eigtime.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
from dmpfold2_b200.engine import Engine
sd, _ = bench.load_weights()
eng = Engine(sd, 0)
for L in (82, 150, 300, 600):
    g = torch.Generator().manual_seed(L)
    x = torch.randn(L, 3, generator=g) * 10
    d = (x[:, None] - x[None]).norm(dim=2) + 0.3 * torch.rand(L, L, generator=g)
    d = (d + d.t()) / 2
    m = 0.5 * (d[0:1, :] ** 2 + d[:, 0:1] ** 2 - d ** 2)
    for _ in range(2):
        vals, vecs = eng.eig_top8(m)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); eng.eig_top8(m); ev1.record(); torch.cuda.synchronize()
    w = torch.linalg.eigvalsh(m.double())[-8:]
    print(L, 'total %.1f us' % (ev0.elapsed_time(ev1) * 1e3), eng.eig_phases(L), 'val err', float((vals.cpu().double() - w).abs().max()), flush=True)
PY

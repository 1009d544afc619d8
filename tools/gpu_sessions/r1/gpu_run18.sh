cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r18_all.log 2>&1
echo "all exit $?" >> gpurun_out/r18_all.log
timeout 300 python - > gpurun_out/r18_eigtime.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
from dmpfold2_b200.engine import Engine
sd, _ = bench.load_weights()
eng = Engine(sd, 0)
for L in (82, 150, 300, 436, 500, 600, 1024, 2048):
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
    w, v = torch.linalg.eigh(m.double())
    from oracle import dmpfold_oracle as O
    v = O.canonical_sign(v)[:, -8:]
    print(L, 'total %.1f us' % (ev0.elapsed_time(ev1) * 1e3), eng.eig_phases(L), 'val err %.2e' % float((vals.cpu().double() - w[-8:]).abs().max()),
          'vec err %.2e' % float((vecs.cpu().double() - v).abs().max()), flush=True)
PY
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r18_bench.json 2> gpurun_out/r18_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv5_tc -s 18 -c 2 -o gpurun_out/r18_conv -f python tools/profile_fold.py 1 f16f8 > gpurun_out/r18_ncu_conv.log 2>&1

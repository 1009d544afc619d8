cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q --timeout 120 -k "tensor_core_gemm_core" > gpurun_out/r2_gemm.log 2>&1
echo "gemm exit $?" >> gpurun_out/r2_gemm.log
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q --timeout 200 -k "conv5 or resblock or resnet_pass" > gpurun_out/r2_conv.log 2>&1
echo "conv exit $?" >> gpurun_out/r2_conv.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 300 > gpurun_out/r2_e2e.log 2>&1
echo "e2e exit $?" >> gpurun_out/r2_e2e.log
for mode in f16x3 f16; do
DMP2_CONV_MODE=$mode timeout 600 python - > gpurun_out/r2_time_$mode.log 2>&1 <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from oracle import dmpfold_oracle as O
from dmpfold2_b200.engine import Engine
sd = O.load_state_dict('dmpfold2_b200/trained_model')
eng = Engine(sd, 0)
base = O.encode_aln(O.read_aln('tests/golden/PF10963.aln'))
for (L, N, n, m) in ((82, 252, 10, 100), (150, 512, 10, 100), (300, 1000, 10, 100)):
    msa = base if L == 82 else O.synth_msa_structured(base, L, N, 0)
    eng.fold_host(msa, None, 0, 0)
    t = time.time(); eng.fold_host(msa, None, n, m); dt = time.time() - t
    print(L, N, n, m, 'wall %.1f ms' % (dt * 1e3), eng.stage_times(), flush=True)
PY
done

cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r1_smi.txt 2>&1
DMP2_CONV_MODE=ffma timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -k "not f16" > gpurun_out/r1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1_pytest.log
DMP2_CONV_MODE=ffma timeout 600 python - > gpurun_out/r1_time.log 2>&1 <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from oracle import dmpfold_oracle as O
from dmpfold2_b200.engine import Engine
sd = O.load_state_dict('dmpfold2_b200/trained_model')
eng = Engine(sd, 0)
base = O.encode_aln(O.read_aln('tests/golden/PF10963.aln'))
for (L, N, n, m) in ((82, 252, 0, 0), (82, 252, 10, 100), (150, 512, 1, 100), (300, 1000, 1, 100)):
    msa = base if L == 82 else O.synth_msa_structured(base, L, N, 0)
    eng.fold_host(msa, None, 0, 0)
    t = time.time(); eng.fold_host(msa, None, n, m); dt = time.time() - t
    print(L, N, n, m, 'wall %.1f ms' % (dt * 1e3), eng.stage_times(), flush=True)
PY

cd $GRAFT_REPO_ROOT
DMP2_VGRU_STAMPS=1 timeout 300 python - > gpurun_out/r28_stamps.log 2>&1 <<'PY'
import os, sys
sys.path.insert(0, '.')
os.environ['DMP2_VGRU'] = 'persist'
import bench, torch
from dmpfold2_b200.engine import Engine
from dmpfold2_b200.predict import read_aln, encode_aln
from dmpfold2_b200.synth import synth_msa_structured
sd, _ = bench.load_weights()
base = encode_aln(read_aln('tests/golden/PF10963.aln'))
eng = Engine(sd, 0)
for (L, N) in ((82, 252), (300, 1000)):
    msa = synth_msa_structured(base, L, N, 0)
    for _ in range(2):
        eng.vgru(msa)
    torch.cuda.synchronize()
PY

cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_stages.py -x -q -k eig > gpurun_out/r34_eig.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r34_eig.log
echo "--- whole-GPU tridiagonalisation for L > cluster capacity" >> gpurun_out/r34_eig.log
timeout 300 python tools/time_eig.py 300 640 700 1024 1500 2048 >> gpurun_out/r34_eig.log 2>&1
echo "--- DMP2_EIG_GRID=0 (16-CTA cluster streaming from L2)" >> gpurun_out/r34_eig.log
DMP2_EIG_GRID=0 timeout 300 python tools/time_eig.py 700 1024 2048 >> gpurun_out/r34_eig.log 2>&1

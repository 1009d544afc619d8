cd $GRAFT_REPO_ROOT
timeout 300 python tools/time_vgru.py > gpurun_out/r27_time_vgru.log 2>&1
echo "time_vgru rc=$?" >> gpurun_out/r27_time_vgru.log
DMP2_VGRU=persist timeout 600 python -m pytest tests -m gpu -x -q -k "vgru or pf10963 or structured" > gpurun_out/r27_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r27_pytest.log
DMP2_VGRU=persist timeout 600 python bench.py > gpurun_out/r27_bench_persist.json 2> gpurun_out/r27_bench_persist.err

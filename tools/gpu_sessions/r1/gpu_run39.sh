cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r39_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r39_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r39_bench.json 2> gpurun_out/r39_bench.err
DMP2_FUSE_STATS=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r39_bench_nofuse.json 2> gpurun_out/r39_bench_nofuse.err

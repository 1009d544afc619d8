cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_r2.py > gpurun_out/r2/18_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r2/18_memcheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 3 python tools/sanitize_r2.py > gpurun_out/r2/18_synccheck.log 2>&1
echo "synccheck exit $?" >> gpurun_out/r2/18_synccheck.log

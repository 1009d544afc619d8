cd $GRAFT_REPO_ROOT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_fold.py > gpurun_out/r19_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r19_memcheck.log
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 3 python tools/sanitize_fold.py > gpurun_out/r19_synccheck.log 2>&1
echo "synccheck exit $?" >> gpurun_out/r19_synccheck.log

cd $GRAFT_REPO_ROOT
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_new.py > gpurun_out/r41_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r41_memcheck.log

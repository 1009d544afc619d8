cd $GRAFT_REPO_ROOT
timeout 100 compute-sanitizer --tool synccheck --error-exitcode 3 python tools/sanitize_new.py > gpurun_out/r42_synccheck.log 2>&1
echo "synccheck exit $?" >> gpurun_out/r42_synccheck.log

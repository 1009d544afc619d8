cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 2400 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_parity_r2.py::test_cfg4_shape_one_pass_and_template_seeded > gpurun_out/r2/08_all.log 2>&1
echo "all exit $?" >> gpurun_out/r2/08_all.log
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r2/08_bench.json 2> gpurun_out/r2/08_bench.err

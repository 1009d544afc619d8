# Last GPU session of round 2 (one B200): validates the state after dmp2_stem / dmp2_head and the CUDA-graph replay of the
# recycling iterations, A/B-times the graph, refreshes the bench line against the re-measured MEASURED_PEAKS.json.
# Ordered by importance: every step has its own timeout and writes its own log.
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_graph
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
# 1. the new tests alone (a failure here must not poison the context of the whole suite)
timeout 300 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_stages.py -m gpu -q -s -k "graph_replay or stem_and_head" > $O/new_tests.log 2>&1
NEW=$?
echo "new tests exit $NEW" >> $O/new_tests.log
# 2. the whole GPU suite in the default configuration (without the new tests if they failed)
if [ $NEW -eq 0 ]; then
  timeout 900 python -m pytest tests -m gpu -q > $O/all_tests.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -q -k "not graph_replay and not stem_and_head" > $O/all_tests.log 2>&1
fi
echo "all exit $?" >> $O/all_tests.log
# 3. every whole-fold test again with the graph replay on for every engine (what flipping the default would run)
DMP2_GRAPH=1 timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_parity_r2.py tests/test_gpu_strip.py -m gpu -q > $O/graph_tests.log 2>&1
echo "graph suite exit $?" >> $O/graph_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke exit $?" >> $O/smoke.log
# 4. the default bench line (full: e2e, roofline, reference cpu_baseline)
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
# 5. A/B of the graph replay: device and host-enqueue time per fold at cfg2 and the cfg3 shape
timeout 300 python tools/time_graph.py 4 > $O/time_graph.log 2>&1
# 6. bench with the graph on, and with the MSA features at high stream priority
DMP2_GRAPH=1 timeout 300 python bench.py --no-cpu-baseline --no-extras > $O/bench_graph.json 2> $O/bench_graph.err
DMP2_SIDE_PRIORITY=high timeout 300 python bench.py --no-cpu-baseline --no-extras > $O/bench_side_high.json 2> $O/bench_side_high.err
# 7. throughput mode at the cfg3 shape on this GPU, eager vs graph replay
timeout 300 python tools/throughput_cfg3.py --targets 64 --streams 1,4 > $O/throughput_eager.log 2>&1
DMP2_GRAPH=1 timeout 300 python tools/throughput_cfg3.py --targets 64 --streams 1,4 > $O/throughput_graph.log 2>&1
ls -la $O > $O/done.txt

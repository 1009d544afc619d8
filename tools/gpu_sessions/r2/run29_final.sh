# final verification of the last code state (eigensolver phases re-cut after tools/gpu_final_r2.sh ran): whole GPU suite, parity
# numbers, smoke, the default bench line, launch list, other configs.  The --set full captures of the conv kernel and the
# reference arm are those of tools/gpu_final_r2.sh (conv_tc.cu and the reference are unchanged since).
cd $GRAFT_REPO_ROOT
O=gpurun_out/final_r2b
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/all_tests.log 2>&1
echo "all exit $?" >> $O/all_tests.log
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -q -s -m gpu > $O/parity.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke exit $?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1
timeout 300 python tools/time_configs.py > $O/configs.log 2>&1

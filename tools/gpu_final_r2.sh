# Round-2 final verification + evidence run (one B200): whole -m gpu suite, smoke(), the default bench line, the reference
# arm, the ncu launch list of the bench command and --set full captures of the dominant kernel (both conv modes) and of
# the covariance-sized tensor-core GEMM.
cd $GRAFT_REPO_ROOT
O=gpurun_out/final_r2
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/all_tests.log 2>&1
echo "all exit $?" >> $O/all_tests.log
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -q -s -m gpu > $O/parity.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke exit $?" >> $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_conv5_tc<\(int\)2" -s 40 -c 2 -f -o $O/conv_f16f8 python tools/time_conv.py f16f8 > $O/ncu_conv_f16f8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_conv5_tc<\(int\)1" -s 60 -c 2 -f -o $O/conv_f16x3 python tools/time_conv.py f16x3 > $O/ncu_conv_f16x3.log 2>&1
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"k_conv5_tc<\(int\)1" -c 10 -f -o $O/gemm_tc python tools/time_conv.py f16f8 > $O/ncu_gemm.log 2>&1
timeout 300 python tools/time_configs.py > $O/configs.log 2>&1

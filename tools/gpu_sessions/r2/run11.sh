cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 600 python tools/throughput_cfg3.py --targets 32 --streams 1,2,3,4,6 > gpurun_out/r2/11_tp_cfg3_1gpu.log 2>&1
timeout 600 python tools/throughput_cfg3.py --targets 32 --streams 3,4 --conv-sms 132 >> gpurun_out/r2/11_tp_cfg3_1gpu.log 2>&1
timeout 600 python tools/throughput_cfg3.py --targets 12 --L 300 --N 1000 --streams 1,2,3,4 > gpurun_out/r2/11_tp_cfg2_1gpu.log 2>&1
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r2/11_bench.json 2> gpurun_out/r2/11_bench.err

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2/01_smi.txt 2>&1
nproc >> gpurun_out/r2/01_smi.txt
timeout 300 python tools/probe_tmem_acc.py > gpurun_out/r2/01_tmem_acc.log 2>&1
timeout 900 python tools/diag_stages64.py 300 1000 0 > gpurun_out/r2/01_stages64.log 2>&1
timeout 600 python tools/diag_fold_dump.py cfg2_s0_n10_m100 300 1000 0 10 100 > gpurun_out/r2/01_fold_cfg2.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/01_bench.json 2> gpurun_out/r2/01_bench.err

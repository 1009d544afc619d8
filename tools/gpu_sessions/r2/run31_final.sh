# Final verification of round 2's last code state (one B200): whole -m gpu suite, smoke(), the default bench line, the ncu launch
# list of the bench command, and --set full captures of the non-conv kernels of a fold (InstanceNorm/gate pass = the HBM-bound
# kernel, stem update, head, eigen step, bi-GRU recurrence, coordinate head, one vgru step).
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_final
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/all_tests.log 2>&1
echo "all exit $?" >> $O/all_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke exit $?" >> $O/smoke.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 400 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"k_norm_gate|k_stem_update|k_in_stats|k_head|k_bigru_rec|k_eig_top8|k_coord_fc|k_gru_operand" -s 2 -c 28 -f -o $O/other_kernels python tools/profile_fold.py 1 f16f8 > $O/ncu_other.log 2>&1
timeout 300 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"k_vgru_step" -s 600 -c 2 -f -o $O/vgru_step python tools/profile_fold.py 0 f16f8 > $O/ncu_vgru.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1
ls -la $O > $O/done.txt

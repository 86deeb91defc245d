# round-end evidence: GPU tests, smoke, both bench arms, ncu launch list and full capture of the top kernels
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/fin_tests.log 2>&1; echo tests rc $?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1; echo smoke rc $?
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/fin_bench.json 2>gpurun_out/fin_bench.err; echo bench rc $?
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > gpurun_out/fin_bench_ref.json 2>gpurun_out/fin_bench_ref.err; echo ref rc $?
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-unpickle --chunks 1"
$CMD > gpurun_out/fin_ncu_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/fin_ncu_launches.csv $CMD > gpurun_out/fin_ncu_launch_run.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_pair|tc_tap_chain2|lbfgs_advance_kernel|energy_grad" -s 90 -c 12 -f -o gpurun_out/fin_prof $CMD > gpurun_out/fin_ncu_full_run.log 2>&1; echo "full rc=$?"
timeout 300 python bench.py --no-cpu-baseline --no-unpickle --steps 5 --warmup 3 --windows 10000 > gpurun_out/fin_sweep_1e4.json 2>gpurun_out/fin_sweep_1e4.err; echo "sweep 1e4 rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-unpickle --steps 2 --warmup 3 --windows 100000 > gpurun_out/fin_sweep_1e5.json 2>gpurun_out/fin_sweep_1e5.err; echo "sweep 1e5 rc=$?"
tail -2 gpurun_out/fin_tests.log

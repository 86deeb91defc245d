timeout 900 python -m pytest tests/test_gpu_lbfgs.py tests/test_gpu_pipeline.py tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/lb2_tests.log 2>&1; echo tests rc $?
B="timeout 300 python bench.py --no-cpu-baseline --no-unpickle --no-e2e --steps 10 --warmup 5"
for v in base generic base generic; do
  if [ $v = base ]; then unset GEM_B200_LIB; else export GEM_B200_LIB=$PWD/globalegomocap_b200/libgem_b200_$v.so; fi
  $B > gpurun_out/lb2_$v.$RANDOM.json 2>>gpurun_out/lb2.err
done
echo done

timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused_energy.py tests/test_gpu_planar.py tests/test_gpu_pipeline.py tests/test_gpu_tc_gemm.py -x -q -m gpu > gpurun_out/en_tests.log 2>&1; echo tests rc $?
B="timeout 300 python bench.py --no-cpu-baseline --no-unpickle --steps 10 --warmup 5"
$B > gpurun_out/en_fixed.json 2>gpurun_out/en_fixed.err
GEM_ENERGY_FIXED=0 $B > gpurun_out/en_dyn.json 2>gpurun_out/en_dyn.err
$B > gpurun_out/en_fixed2.json 2>gpurun_out/en_fixed2.err
GEM_ENERGY_FIXED=0 $B > gpurun_out/en_dyn2.json 2>gpurun_out/en_dyn2.err
echo done

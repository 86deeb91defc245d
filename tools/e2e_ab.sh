timeout 600 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_kernels.py tests/test_gpu_fused_energy.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/chain_tests.log 2>&1; echo tests rc $?
timeout 120 python tools/dbg_chain.py 468 > gpurun_out/dbg_chain_468_stream.txt 2>&1
GEM_CHAIN_STREAM=0 timeout 120 python tools/dbg_chain.py 468 > gpurun_out/dbg_chain_468_nostream.txt 2>&1
B="timeout 300 python bench.py --no-cpu-baseline --no-unpickle --no-e2e --steps 10 --warmup 5"
$B > gpurun_out/chain_stream.json 2>gpurun_out/chain_stream.err
GEM_CHAIN_STREAM=0 $B > gpurun_out/chain_nostream.json 2>gpurun_out/chain_nostream.err
$B > gpurun_out/chain_stream2.json 2>gpurun_out/chain_stream2.err
GEM_CHAIN_STREAM=0 $B > gpurun_out/chain_nostream2.json 2>gpurun_out/chain_nostream2.err
echo done

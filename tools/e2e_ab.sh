timeout 600 python -m pytest tests/test_gpu_planar.py -q -m gpu > gpurun_out/probe_tests.log 2>&1; echo tests rc $?
B="timeout 300 python bench.py --no-cpu-baseline --no-unpickle --steps 10 --warmup 5 --heat-layout tiled"
run() { name=$1; shift; env "$@" $B > gpurun_out/e2e_t_$name.json 2>gpurun_out/e2e_t_$name.err; }
run pf8 X=1
run pf4 GEM_TEXEL_PREFETCH_CTAS=4
run pf16 GEM_TEXEL_PREFETCH_CTAS=16
run pf8c16 GEM_TEXEL_COLD_CTAS=16
run pf8c64 GEM_TEXEL_COLD_CTAS=64
run pf8t256 GEM_TEXEL_PREFETCH_THREADS=256
echo done

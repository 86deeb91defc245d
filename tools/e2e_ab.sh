B="python bench.py --no-cpu-baseline --no-unpickle --steps 10 --warmup 5"
run() { name=$1; shift; env "$@" $B > gpurun_out/e2e_cold_$name.json 2>gpurun_out/e2e_cold_$name.err; }
run pf8 GEM_TEXEL_PREFETCH_CTAS=8
run pf8b GEM_TEXEL_PREFETCH_CTAS=8
run pf8r4 GEM_TEXEL_PREFETCH_CTAS=8 GEM_TEXEL_COLD_ROWS=4
run pf16 GEM_TEXEL_PREFETCH_CTAS=16
run pf8nochain GEM_TEXEL_PREFETCH_CTAS=8 GEM_TEXEL_COLD_CHAIN=0
run pf16t256 GEM_TEXEL_PREFETCH_CTAS=16 GEM_TEXEL_PREFETCH_THREADS=256
run pf8c64 GEM_TEXEL_PREFETCH_CTAS=8 GEM_TEXEL_COLD_CTAS=64
echo done

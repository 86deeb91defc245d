B="timeout 300 python bench.py --no-cpu-baseline --no-unpickle --no-e2e --steps 10 --warmup 5"
for v in 0 148 100 0 148 128; do
  GEM_LBFGS_SMS=$v $B > gpurun_out/lb3_sms$v.$RANDOM.json 2>>gpurun_out/lb3.err
done
echo done

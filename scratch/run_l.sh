for st in 0 3000 6000 12000 24000; do
for w in C2 C3; do
  PMOC_STAGGER=$st python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench_$1_$w.json 2> gpurun_out/bench_$1_$w.err
  python -c "import json;d=json.load(open('gpurun_out/bench_$1_$w.json'));print('stagger $st $w',d['value'],d['roofline']['frac'])"; tail -3 gpurun_out/bench_$1_$w.err
done; done
python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('default C2',d['value'],d['roofline']['frac'])"

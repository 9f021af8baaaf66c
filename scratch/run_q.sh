for wpb in 15 14 13 12 7 5 4; do
  PMOC_WPB=$wpb python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
  python -c "import json;d=json.load(open('gpurun_out/bench_tmp.json'));print('C4 wpb $wpb',d['value'],d['roofline']['frac'])"; tail -2 gpurun_out/bench_tmp.err
done
for wpb in 16 13 8 5 4; do
  PMOC_WPB=$wpb python bench.py --workload C3 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
  python -c "import json;d=json.load(open('gpurun_out/bench_tmp.json'));print('C3 wpb $wpb',d['value'],d['roofline']['frac'])"; tail -2 gpurun_out/bench_tmp.err
done
for wpb in 8 4 2; do
  PMOC_WPB=$wpb python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
  python -c "import json;d=json.load(open('gpurun_out/bench_tmp.json'));print('C2 wpb $wpb',d['value'],d['roofline']['frac'])"; tail -2 gpurun_out/bench_tmp.err
done

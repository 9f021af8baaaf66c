tag=$1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}_default.json 2> gpurun_out/bench_${tag}_default.err; tail -c 300 gpurun_out/bench_${tag}_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err; tail -c 300 gpurun_out/bench_${tag}_reference.json
for w in C3 C4 C5_4096 twobasin C1; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_${tag}_$w.json 2> gpurun_out/bench_${tag}_$w.err
  python -c "import json;d=json.load(open('gpurun_out/bench_${tag}_$w.json'));print('$w',d['value'],d['roofline']['frac'],d['e2e']['value'] if d['e2e'] else None)"; tail -3 gpurun_out/bench_${tag}_$w.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_list_${tag}.log 2>&1
bash tools/prof_all.sh $tag "C2 65536 720 k_model 2" "C3 32768 240 k_model 2 2400" "C4 32768 240 k_model 1 4800" "C5_4096 2048 200 k_wide_steps2 1"

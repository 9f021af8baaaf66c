tag=$1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${tag}_default.json 2> gpurun_out/bench_${tag}_default.err; tail -c 400 gpurun_out/bench_${tag}_default.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err; tail -c 600 gpurun_out/bench_${tag}_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_list_${tag}.log 2>&1
bash scratch/prof_all.sh $tag "C2 65536 720 k_model 2" "C3 32768 240 k_model 2 2400" "C4 32768 240 k_model 1 4800" "C5_4096 2048 200 k_wide_steps2 1"

set -x
mkdir -p gpurun_out
export SVS_BENCH_FRAMES=64
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_plain64.json 2>gpurun_out/r2t.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2t_ncu1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:blk -c 2 -o gpurun_out/r2_final python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2t_ncu2.log 2>&1
tail -2 gpurun_out/r2t_ncu2.log
unset SVS_BENCH_FRAMES
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_1.json 2>gpurun_out/r2t_b1.err; tail -c 300 gpurun_out/r2_bench_1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>gpurun_out/r2t_ref.err; tail -c 200 gpurun_out/r2_bench_reference.json
timeout 600 python bench.py --workload 4k --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_4k_1_ac63.json 2>gpurun_out/r2t_4k.err
timeout 600 python bench.py --workload 4k --num-ac 10 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_4k_1_ac10.json 2>>gpurun_out/r2t_4k.err
timeout 600 python profiles/n2_throughput.py > gpurun_out/r2_n2_throughput.txt 2>gpurun_out/r2t_n2.err; cat gpurun_out/r2_n2_throughput.txt
timeout 1200 python profiles/sweep.py > gpurun_out/r2_sweep.jsonl 2>gpurun_out/r2t_sweep.err; wc -l gpurun_out/r2_sweep.jsonl

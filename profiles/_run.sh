set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -4 gpurun_out/r2j_pytest.log
SVS_TEST_LIB=variants/libsvs_variants.so timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2j_pytest_variants.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_variants.log
tail -3 gpurun_out/r2j_pytest_variants.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; tail -2 gpurun_out/r2j_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 3000 gpurun_out/r2j_bench.json; tail -3 gpurun_out/r2j_bench.err
timeout 600 python bench.py --workload 4k --steps 5 --warmup 3 --no-cpu > gpurun_out/r2j_bench_4k.json 2> gpurun_out/r2j_bench_4k.err; tail -c 1500 gpurun_out/r2j_bench_4k.json; tail -3 gpurun_out/r2j_bench_4k.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2j_ref.json 2> gpurun_out/r2j_ref.err; tail -c 1200 gpurun_out/r2j_ref.json

set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
tail -3 gpurun_out/r2r_pytest.log
timeout 300 python profiles/ab_kernels.py --frames 600 --families 5 --ac 63,10 --tag fix2 > gpurun_out/r2r_ab.jsonl 2>gpurun_out/r2r_ab.err
timeout 300 python profiles/ab_kernels.py --frames 600 --families 5 --ac 63 --tag fix2_sse --sse >> gpurun_out/r2r_ab.jsonl 2>>gpurun_out/r2r_ab.err
cat gpurun_out/r2r_ab.jsonl

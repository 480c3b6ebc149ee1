set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:embed_blk -c 1 -o gpurun_out/r2n_side python profiles/ab_kernels.py --frames 64 --families 5 --ac 63 --iters 1 --sse > gpurun_out/r2n_ncu.log 2>&1
tail -2 gpurun_out/r2n_ncu.log

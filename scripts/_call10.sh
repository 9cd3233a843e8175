nvidia-smi --query-gpu=index,name --format=csv,noheader | head -3
python -m pytest tests/test_sharded_gpu.py -x -q > gpurun_out/r02_call10_sharded_nccl.log 2>&1; tail -4 gpurun_out/r02_call10_sharded_nccl.log
VSTAB_BENCH_PHASES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
cat gpurun_out/r02_bench_n2.json; grep phases gpurun_out/r02_bench_n2.err | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/cfg5_scale.py --frames-per-gpu 250 --steps 3 --warmup 2 --e2e-frames 32 > gpurun_out/r02_cfg5_n2.json 2> gpurun_out/r02_cfg5_n2.err
cat gpurun_out/r02_cfg5_n2.json; tail -3 gpurun_out/r02_cfg5_n2.err

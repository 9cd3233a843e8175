nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" | head -8
cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null | head -2
VSTAB_BENCH_PHASES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
cat gpurun_out/r02_bench_n8.json | cut -c1-3000; grep phases gpurun_out/r02_bench_n8.err | tail -8
for n in 8 4 2 1; do
  if [ $n = 1 ]; then python scripts/cfg5_scale.py --frames-per-gpu 250 --steps 3 --warmup 2 --e2e-frames 32 > gpurun_out/r02_cfg5_scale_n$n.json 2> gpurun_out/r02_cfg5_scale_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n scripts/cfg5_scale.py --frames-per-gpu 250 --steps 3 --warmup 2 --e2e-frames 32 > gpurun_out/r02_cfg5_scale_n$n.json 2> gpurun_out/r02_cfg5_scale_n$n.err; fi
  tail -1 gpurun_out/r02_cfg5_scale_n$n.json | cut -c1-1500
done

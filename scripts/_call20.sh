VSTAB_BENCH_PHASES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e > gpurun_out/r02_bench_n8c.json 2> gpurun_out/r02_bench_n8c.err
cut -c1-330 gpurun_out/r02_bench_n8c.json; grep phases gpurun_out/r02_bench_n8c.err | tail -2 | cut -c1-1700
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e 2>/dev/null | cut -c1-330
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e 2>/dev/null | cut -c1-330
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | cut -c1-330
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2958$n scripts/cfg5_scale.py --frames-per-gpu 250 --steps 3 --warmup 2 --e2e-frames 0 > gpurun_out/r02_cfg5b_scale_n$n.json 2> /dev/null
python -c "import json; d=json.loads(open('gpurun_out/r02_cfg5b_scale_n$n.json').read().strip().splitlines()[-1]); r=d['device_resident']; print($n, r['ms_per_step'], r['frames_per_s']); print(r['gpu_ms_between_marks_per_rank'])"
done
python scripts/cfg5_scale.py --frames-per-gpu 250 --steps 3 --warmup 2 --e2e-frames 0 > gpurun_out/r02_cfg5b_scale_n1.json 2> /dev/null
python -c "import json; d=json.loads(open('gpurun_out/r02_cfg5b_scale_n1.json').read().strip().splitlines()[-1]); r=d['device_resident']; print(1, r['ms_per_step'], r['frames_per_s'])"

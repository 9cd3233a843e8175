python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests_final.txt 2>&1; tail -4 gpurun_out/r02_gpu_tests_final.txt
python -c 'import __graft_entry__ as g; g.smoke()' >> gpurun_out/r02_gpu_tests_final.txt 2>&1; tail -1 gpurun_out/r02_gpu_tests_final.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; cut -c1-300 gpurun_out/r02_bench_final.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_final.json 2>&1; cut -c1-200 gpurun_out/r02_bench_ref_final.json
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_flow1080p.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_launches_flow1080p.csv | head -24
python scripts/kernel_probe.py > gpurun_out/r02_kernel_probe.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:warp_stream -s 3 -c 1 -o gpurun_out/r02_warp_stream python scripts/kernel_probe.py > /dev/null 2>&1
python scripts/dis_profile.py 3 > gpurun_out/r02_dis_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"vr_fused_kernel|patch_search_kernel" -s 28 -c 4 -o gpurun_out/r02_dis_final python scripts/dis_profile.py 1 > /dev/null 2>&1
cat gpurun_out/r02_dis_plain3.log gpurun_out/r02_kernel_probe.log | cut -c1-600

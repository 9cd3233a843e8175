python -m pytest tests/test_dis_gpu.py tests/test_small_frames_gpu.py tests/test_warp_gpu.py tests/test_flow_gpu.py -x -q > gpurun_out/r02_call7_tests.log 2>&1
tail -5 gpurun_out/r02_call7_tests.log
export SWEEP_CONFIGS='[{"VSTAB_DIS_GROUPS":1},{"VSTAB_DIS_GROUPS":2},{"VSTAB_DIS_GROUPS":1,"VSTAB_PS_WPC":2},{"VSTAB_DIS_GROUPS":2,"VSTAB_PS_WPC":2},{"VSTAB_DIS_GROUPS":3,"VSTAB_PS_WPC":2},{"VSTAB_DIS_GROUPS":1,"VSTAB_PS_WPC":8},{"VSTAB_DIS_GROUPS":1,"VSTAB_PS_NOPACK":1,"VSTAB_PS_WPC":8}]'
python scripts/dis_sweep.py
VSTAB_DIS_GROUPS=1 python scripts/dis_profile.py 3 > gpurun_out/r02_dis_plain2.log 2>&1 && VSTAB_DIS_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:"patch_search_kernel" -s 7 -c 1 -o gpurun_out/r02_ps_v2 python scripts/dis_profile.py 1 > gpurun_out/r02_ps_ncu.log 2>&1
cat gpurun_out/r02_dis_plain2.log

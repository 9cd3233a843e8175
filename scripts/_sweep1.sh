python -m pytest tests/test_dis_gpu.py tests/test_small_frames_gpu.py tests/test_warp_gpu.py tests/test_warp_stream_gpu.py tests/test_flow_gpu.py tests/test_streamed_gpu.py -x -q > gpurun_out/r02_call2_tests.log 2>&1
tail -15 gpurun_out/r02_call2_tests.log
export SWEEP_CONFIGS='[{"VSTAB_PS_NOPACK":1,"VSTAB_PS_WPC":8,"VSTAB_VR_CLUSTER":8},{"VSTAB_PS_WPC":8,"VSTAB_VR_CLUSTER":8},{"VSTAB_PS_WPC":4,"VSTAB_VR_CLUSTER":8},{"VSTAB_PS_WPC":2,"VSTAB_VR_CLUSTER":8},{"VSTAB_PS_WPC":1,"VSTAB_VR_CLUSTER":8},{"VSTAB_PS_WPC":2},{"VSTAB_PS_WPC":2,"VSTAB_VR_CLUSTER":6},{"VSTAB_PS_WPC":2,"VSTAB_VR_CLUSTER":7},{"VSTAB_PS_WPC":2,"VSTAB_VR_CLUSTER":4},{"VSTAB_PS_WPC":2,"VSTAB_DIS_GROUPS":2},{"VSTAB_PS_WPC":2,"VSTAB_DIS_GROUPS":3},{"VSTAB_PS_WPC":2,"VSTAB_DIS_GROUPS":4},{"VSTAB_PS_WPC":2,"VSTAB_DIS_GROUPS":2,"VSTAB_VR_CLUSTER":8},{"VSTAB_PS_WPC":2,"VSTAB_DIS_GROUPS":4,"VSTAB_VR_CLUSTER":8},{"VSTAB_PS_WPC":4,"VSTAB_DIS_GROUPS":4,"VSTAB_VR_CLUSTER":8}]'
python scripts/dis_sweep.py > gpurun_out/r02_sweep1_m5.log 2>&1
export SWEEP_CONFIGS='[{"VSTAB_PS_WPC":2},{"VSTAB_PS_WPC":2,"VSTAB_VR_CLUSTER":7},{"VSTAB_PS_WPC":2,"VSTAB_VR_CLUSTER":8},{"VSTAB_PS_WPC":2,"VSTAB_DIS_GROUPS":2},{"VSTAB_PS_WPC":2,"VSTAB_DIS_GROUPS":4},{"VSTAB_PS_WPC":2,"VSTAB_DIS_GROUPS":4,"VSTAB_VR_CLUSTER":8}]'
VSTAB_LIB=$PWD/build_ab/libvstab_m6.so python scripts/dis_sweep.py > gpurun_out/r02_sweep1_m6.log 2>&1
VSTAB_LIB=$PWD/build_ab/libvstab_m8.so python scripts/dis_sweep.py > gpurun_out/r02_sweep1_m8.log 2>&1
cat gpurun_out/r02_sweep1_m5.log gpurun_out/r02_sweep1_m6.log gpurun_out/r02_sweep1_m8.log

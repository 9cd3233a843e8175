export SWEEP_CONFIGS='[{"VSTAB_DIS_GROUPS":1},{"VSTAB_DIS_GROUPS":2},{"VSTAB_DIS_GROUPS":2,"VSTAB_PS_WPC":8},{"VSTAB_DIS_GROUPS":3}]'
python scripts/dis_sweep.py
VSTAB_LIB=$PWD/build_ab/libvstab_ps2.so python scripts/dis_sweep.py
python scripts/cfg5_scale.py --frames-per-gpu 250 --steps 3 --warmup 2 --e2e-frames 32 --cpu-frames 6 > gpurun_out/r02_cfg5_n1.json 2> gpurun_out/r02_cfg5_n1.err
cat gpurun_out/r02_cfg5_n1.json; tail -3 gpurun_out/r02_cfg5_n1.err

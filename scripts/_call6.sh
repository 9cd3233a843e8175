python -m pytest tests/test_dis_gpu.py tests/test_small_frames_gpu.py tests/test_warp_gpu.py tests/test_flow_gpu.py -x -q > gpurun_out/r02_call6_tests.log 2>&1
tail -5 gpurun_out/r02_call6_tests.log
export SWEEP_CONFIGS='[{"VSTAB_DIS_GROUPS":1},{"VSTAB_DIS_GROUPS":2},{"VSTAB_DIS_GROUPS":1,"VSTAB_PS_WPC":2},{"VSTAB_DIS_GROUPS":2,"VSTAB_PS_WPC":2},{"VSTAB_DIS_GROUPS":3,"VSTAB_PS_WPC":2},{"VSTAB_DIS_GROUPS":1,"VSTAB_PS_NOPACK":1,"VSTAB_PS_WPC":8}]'
python scripts/dis_sweep.py
python scripts/phase_probe.py > gpurun_out/r02_phase_probe1.log 2>&1; head -24 gpurun_out/r02_phase_probe1.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_dis_launches.csv python scripts/dis_profile.py 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r02_dis_launches.csv 2>/dev/null | head -30
python - <<'PY'
import csv,collections
rows=[]
for row in csv.DictReader([l for l in open("gpurun_out/r02_dis_launches.csv") if not l.startswith("==")]):
    if row.get("Metric Name")=="gpu__time_duration.sum": rows.append((row["Kernel Name"], float(row["Metric Value"].replace(",",""))))
# last DIS call = everything after the last zero_kernel pair start
idx=[i for i,r in enumerate(rows) if "tensor_rows" in r[0]]
last=rows[idx[-4]-2:] if len(idx)>=4 else rows
for n,t in last: print("%-60s %9.1f us"%(n[:60], t/1e3))
PY

#!/bin/bash
# Round artefacts that need ncu (run on the GPU box):   bash profiles/tools/capture.sh r02
# launch list of a short bench run + `ncu --set full` of the C2 launch, the throughput instantiation on 4096 x 20000
# astrometry / RV+jitter, and the trajectory-resident explorer; raw / source pages exported as CSV into gpurun_out/
R=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_kepler_like -s 3 -c 1 -o /tmp/${R}_c2 python profiles/tools/prof_c2.py 2>&1 | tail -1
ncu -i /tmp/${R}_c2.ncu-rep --page raw --csv > gpurun_out/${R}_c2_raw.csv; ncu -i /tmp/${R}_c2.ncu-rep --page source --csv > gpurun_out/${R}_c2_src.csv
ncu --set full --clock-control none --import-source on -k regex:k_kepler_like -s 1 -c 1 -o /tmp/${R}_astrom python profiles/tools/prof_kinds.py 4096 20000 2 2>&1 | tail -1
ncu -i /tmp/${R}_astrom.ncu-rep --page raw --csv > gpurun_out/${R}_astrom_raw.csv; ncu -i /tmp/${R}_astrom.ncu-rep --page source --csv > gpurun_out/${R}_astrom_src.csv
ncu --set full --clock-control none -k regex:k_kepler_like -s 3 -c 1 -o /tmp/${R}_rv python profiles/tools/prof_kinds.py 4096 20000 2 2>&1 | tail -1
ncu -i /tmp/${R}_rv.ncu-rep --page raw --csv > gpurun_out/${R}_rv_raw.csv
ncu --set full --clock-control none --import-source on -k regex:k_hmc_resident -s 1 -c 1 -o /tmp/${R}_resident python profiles/tools/hmc_time.py 1024 10 10 2>&1 | tail -1
ncu -i /tmp/${R}_resident.ncu-rep --page raw --csv > gpurun_out/${R}_resident_raw.csv; ncu -i /tmp/${R}_resident.ncu-rep --page source --csv > gpurun_out/${R}_resident_src.csv
ls -la gpurun_out

"""Summarise an `ncu --page raw --csv` dump: duration, occupancy, pipe utilisation, top stall reasons.
    ncu_summary.py raw.csv [--json out.json]     (--json: the first kernel's key numbers, read by bench.py for `roofline.traffic`)"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
for r in rows[2:]:
    d = dict(zip(h, r))
    print('----', d['Kernel Name'][:60], 'grid', d['Grid Size'], 'block', d['Block Size'])
    for k in keys:
        if k in d:
            print('  %-70s %s' % (k, d[k]))
    st = []
    for k in h:
        if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio'):
            try:
                st.append((k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''),
                           float(d[k].replace(',', ''))))
            except ValueError:
                pass
    print('  stalls (warps per issue-active cycle):', ', '.join('%s=%.2f' % kv for kv in sorted(st, key=lambda x: -x[1])[:8]))
    for k in h:
        if 'pipe' in k and k.endswith('pct_of_peak_sustained_active'):
            try:
                v = float(d[k].replace(',', ''))
            except ValueError:
                continue
            if v > 4:
                print('  PIPE %-64s %.1f' % (k, v))

if "--json" in sys.argv:
    d = dict(zip(h, rows[2]))
    num = lambda k: float(d[k].replace(",", "")) if d.get(k) not in (None, "", "n/a") else None
    unit = dict(zip(h, rows[1]))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {"kernel": d["Kernel Name"], "grid": d["Grid Size"], "block": d["Block Size"],
           "duration_us": num("gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(unit.get("gpu__time_duration.sum", "us"), 1.0),
           "dram_bytes_read": num("dram__bytes_read.sum") * scale.get(unit.get("dram__bytes_read.sum", "byte"), 1.0),
           "dram_bytes_write": num("dram__bytes_write.sum") * scale.get(unit.get("dram__bytes_write.sum", "byte"), 1.0),
           "registers": num("launch__registers_per_thread"),
           "fp64_pipe_pct": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
           "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active")}
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)

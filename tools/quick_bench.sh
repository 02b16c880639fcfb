#!/bin/bash
# quick A/B on the GPU box: tools/quick_bench.sh [bench args]
python bench.py --no-cpu-baseline "$@" | python -c "
import json,sys
d=json.load(sys.stdin)
print(json.dumps({'value':round(d['value']/1e9,3),'ms_step':round(d['ms_per_step'],3),'e2e_ms':round(d['e2e']['ms_per_step'],3),'solve':{k:(round(v,3) if isinstance(v,float) else v) for k,v in d['solve_ms_per_step'].items()},'roofline_frac':round(d['roofline']['frac'],3),'active':d['roofline']['active_joint_iterations'],'launches':d['gpu_launches'],'stages':d.get('resident_stage_wall_ms'),'stages_max':d.get('resident_stage_wall_ms_max'),'allocs':d.get('device_allocations_in_timed_region'),'clocks':d['clocks']}))"

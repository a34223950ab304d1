#!/bin/bash
# round-end evidence run (1 GPU): parity suite, smoke, bench (both arms); add "ncu" to also capture the launch list of the bench command
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
echo "== GPU suite"; timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f launches %d frac %.4f'%(d['value'],d['e2e']['value'],d['gpu_launches'],d['roofline']['frac']), d['roofline']['solo_stage_ms'])
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2> gpurun_out/bench_r1_reference.err; tail -c 200 gpurun_out/bench_r1_reference.json; echo
if [ "$1" = "ncu" ]; then
  echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/ncu_launches_r1.csv python bench.py --steps 1 --warmup 1 --streams 4 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-160
fi

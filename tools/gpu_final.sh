#!/bin/bash
# round-end evidence run (1 GPU): parity suite, smoke, bench (both arms, every config), ncu launch list of the bench command and one
# `--set full` capture of a short 4K GOP.  Everything lands in gpurun_out/ (copy what should be judged into profiles/).
# usage: tools/gpu_final.sh [tag] [skip-tests]
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=${1:-r2}
mkdir -p gpurun_out
if [ "$2" != "skip-tests" ]; then
  echo "== GPU suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
fi
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for cfg in 4k 1080p 4k_slow_crf 8k; do
  echo "== bench $cfg"; timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_$cfg.json 2> gpurun_out/bench_${TAG}_$cfg.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${TAG}_$cfg.json').read().strip().splitlines()[-1])
    print('value %.0f e2e %.0f cli %s launches %d dominant %s frac %.5f'%(d['value'],d['e2e']['value'],(d.get('e2e_cli') or {}).get('value'),d['gpu_launches'],d['roofline']['kernel'],d['roofline']['frac']), d['quality'].get('kbps'), d['quality'].get('psnr_y'))
    print({k:round(v,4) for k,v in d['roofline']['solo_stage_ms'].items()})
except Exception as e: print('failed', e)
PY
  echo "== reference arm $cfg"; timeout 900 python bench.py --impl reference --config $cfg --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_${cfg}_reference.json 2> gpurun_out/bench_${TAG}_${cfg}_reference.err; tail -c 300 gpurun_out/bench_${TAG}_${cfg}_reference.json; echo
done
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 600 --csv --log-file gpurun_out/ncu_launches_${TAG}.csv python bench.py --steps 1 --warmup 1 --streams 4 --no-cli > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-160
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_${TAG}_final python tools/profile_driver.py 3 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -3

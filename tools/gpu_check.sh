#!/bin/bash
# development helper: full GPU parity suite, then bench under the tuning knobs named on the command line (e.g. "KS_INTRA_MINB=1" "KS_RECON_MINB=5")
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
echo "== GPU suite"; timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
i=0
for knob in "" "$@"; do
  echo "== bench [$knob]"
  env $knob timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_knob_$i.json 2> gpurun_out/bench_knob_$i.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_knob_$i.json').read().strip().splitlines()[-1])
print('value %.0f e2e %.0f'%(d['value'],d['e2e']['value']), {k:round(v,4) for k,v in d['roofline']['solo_stage_ms'].items()})
PY
  i=$((i+1))
done

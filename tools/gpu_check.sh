#!/bin/bash
# development helper: full GPU parity suite, bench (TMA on / off), one ncu --set full capture of the three big kernels
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
echo "== GPU suite"; timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -6; rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then
  echo "== suite with SAO TMA staging off"; KS_TMA_MASK=0 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
fi
timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err; tail -c 300 gpurun_out/bench_r1_g.json; echo
KS_TMA_MASK=0 timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_f.json 2> gpurun_out/bench_r1_f.err; tail -c 300 gpurun_out/bench_r1_f.json; echo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ks_me_kernel|ks_recon_inter|ks_sao" -s 3 -c 3 -f -o gpurun_out/prof_r1_h python tools/profile_driver.py 4 > gpurun_out/ncu_full_h.log 2>&1; tail -2 gpurun_out/ncu_full_h.log

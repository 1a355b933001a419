#!/bin/bash
# 16 -> 16 sliding kernel: two CTAs per SM (115 registers) against three (96 registers + 8 bytes of spill)
mkdir -p gpurun_out
for m in 0 1 0 1; do
  echo "== MINB3=$m" >> gpurun_out/f5_minb.txt
  B200SEG_SLIDE_MINB3=$m timeout 100 python scripts/kbench.py head l0 >> gpurun_out/f5_minb.txt 2>&1
  B200SEG_SLIDE_MINB3=$m timeout 120 python scripts/fuse_bench.py >> gpurun_out/f5_minb.txt 2>&1
  B200SEG_SLIDE_MINB3=$m timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-roofline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step ms', d['ms_per_step'])" >> gpurun_out/f5_minb.txt
done
cat gpurun_out/f5_minb.txt

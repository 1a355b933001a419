#!/bin/bash
# Where does the sliding kernel's time go?  head layer (2 x 128^3), parts of the kernel switched off (timings only).
mkdir -p gpurun_out
for m in 0 1; do for dbg in 0 1 2 3 4 5 6 7; do
  echo "== MINB3=$m DEBUG=$dbg" >> gpurun_out/f9_slidedbg.txt
  B200SEG_SLIDE_MINB3=$m B200SEG_SLIDE_DEBUG=$dbg timeout 60 python scripts/kbench.py head 2>&1 | grep -E "fprop" | cut -c1-120 >> gpurun_out/f9_slidedbg.txt
done; done
cat gpurun_out/f9_slidedbg.txt

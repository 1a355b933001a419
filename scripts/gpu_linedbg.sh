#!/bin/bash
# Where does the line kernel's time go?  head layer (2 x 128^3) fprop with parts of the kernel switched off.
mkdir -p gpurun_out
for eg in 1 2; do for dbg in 0 1 2 3 4 5 6 7; do
  echo "== EG=$eg DEBUG=$dbg" >> gpurun_out/f3_linedbg.txt
  B200SEG_LINE_CONV=1 B200SEG_LINE_W128=1 B200SEG_LINE_EG=$eg B200SEG_LINE_DEBUG=$dbg timeout 60 python scripts/head_layer.py 2>&1 | grep -E "fprop|dgrad" >> gpurun_out/f3_linedbg.txt
done; done
cat gpurun_out/f3_linedbg.txt
